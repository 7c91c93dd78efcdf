// kmc_main.cpp -- host driver with the call order of the reference's main() (src/kmc_main.cpp:56-603) for the
// field-solve + event-selection path, written against include/gpu_solvers_b200.hpp (reference entry-point names) and
// libkmc_b200.so.  Reads parameters.txt and the xyz structure files unchanged; writes output<size>_<rank>.txt with the
// reference's "KMC time is:" lines and Results_<V>/snapshot_*.xyz in the reference's snapshot format
// (src/Device.cpp:214-232), so a run can be diffed against structures/5nm_device/expected_output.
//
//   usage: kmc_b200_run parameters.txt [max_supersteps]
#include <sys/stat.h>

#include <algorithm>
#include <chrono>
#include <fstream>
#include <iostream>
#include <sstream>

#include "../../include/gpu_solvers_b200.hpp"

static const char *kElementNames[] = {"d", "Od", "V", "O", "Hf", "Ni", "Ti", "Pt", "N"};

struct HostDevice {  // what Device holds on the host (src/Device.h)
    int N = 0;
    std::vector<int> site_element, site_charge, site_layer;
    std::vector<double> site_x, site_y, site_z, site_potential_charge, site_power;
};

static void write_snapshot(const HostDevice &d, const std::string &folder, const std::string &file) {
    std::ofstream fout(("./" + folder + "/" + file).c_str());
    fout << d.N << "\n\n";
    for (int i = 0; i < d.N; i++)
        fout << kElementNames[d.site_element[i]] << "   " << d.site_x[i] << "   " << d.site_y[i] << "   " << d.site_z[i]
             << "   " << d.site_potential_charge[i] << "   " << d.site_power[i] << "\n";
}

int main(int argc, char **argv) {
    if (argc < 2) { std::fprintf(stderr, "usage: %s parameters.txt [max_supersteps]\n", argv[0]); return 2; }
    const long max_steps = argc > 2 ? std::atol(argv[2]) : -1;
    kmcb200_comm_t world;  // one process per GPU; this driver runs rank 0 of 1
    kmcb200_params p;
    KMCB200_CHECK(kmcb200_parse_parameters(argv[1], &p));
    std::string dir(argv[1]);
    dir = dir.find_last_of('/') == std::string::npos ? "." : dir.substr(0, dir.find_last_of('/'));
    std::ofstream outputFile("output" + std::to_string(world.size) + "_" + std::to_string(world.rank) + ".txt");
    std::ostringstream outputBuffer;

    // ---- Device (src/kmc_main.cpp:127-155) ------------------------------------------------------------------
    std::vector<std::string> xyz_files;
    if (p.restart) {
        outputBuffer << "Restarting from " << p.restart_xyz_file << "\n";
        xyz_files.push_back(p.restart_xyz_file);
    } else {
        xyz_files.push_back(p.atom_xyz_file);
        xyz_files.push_back(p.interstitial_xyz_file);
    }
    HostDevice device;
    for (auto &f : xyz_files) {
        std::string path = dir + "/" + f;
        int n = kmcb200_xyz_count(path.c_str());
        if (n < 0) { std::fprintf(stderr, "%s\n", kmcb200_last_error()); return 1; }
        size_t o = device.site_element.size();
        device.site_element.resize(o + n); device.site_x.resize(o + n); device.site_y.resize(o + n); device.site_z.resize(o + n);
        if (kmcb200_read_xyz(path.c_str(), n, device.site_element.data() + o, device.site_x.data() + o,
                             device.site_y.data() + o, device.site_z.data() + o) < 0) {
            std::fprintf(stderr, "%s\n", kmcb200_last_error());
            return 1;
        }
    }
    device.N = (int)device.site_element.size();
    if (p.pristine) {
        int nv = kmcb200_make_substoichiometric(device.N, device.site_element.data(), p.initial_vacancy_concentration, p.rnd_seed);
        std::cout << nv << " oxygen atoms will be converted to vacancies" << std::endl;
    }
    device.site_charge.assign(device.N, 0);
    device.site_potential_charge.assign(device.N, 0.0);
    device.site_power.assign(device.N, 0.0);
    device.site_layer.resize(device.N);
    KMCB200_CHECK(kmcb200_assign_layers(device.N, device.site_x.data(), device.site_layer.data()));
    std::cout << "Loaded " << device.N << " sites into device\n";

    // ---- communicator layout, KMC process, GPU buffers (src/kmc_main.cpp:161-191) ----------------------------------
    const int NL = p.num_atoms_first_layer;
    KMC_comm kmc_comm(world, device.N - 2 * NL, 0, device.N, device.N);
    RandomNumberGenerator random_generator;
    random_generator.setSeed(1);  // rnd_seed_kmc, src/structure_input.h:8
    double E_gen[5], E_rec[5], E_Vdiff[5], E_Odiff[5];
    kmcb200_layer_table(E_gen, E_rec, E_Vdiff, E_Odiff, nullptr, nullptr);
    kmcb200_ctx *ctx = nullptr;
    KMCB200_CHECK(kmcb200_create(&ctx, 0, nullptr));
    std::vector<int> metals(p.metals, p.metals + p.num_metals);
    GPUBuffers gpubuf(ctx, device.site_layer, p.freq, device.N, device.site_element, device.site_x, device.site_y,
                      device.site_z, 52, p.sigma, p.k, p.lattice, metals, p.background_temp);

    // ---- neighbour lists + K sparsity (src/kmc_main.cpp:197-218) ----------------------------------------------------
    compute_neighbor_list(kmc_comm.comm_events, kmc_comm.counts_events.data(), kmc_comm.displs_events.data(), gpubuf);
    if (p.solve_potential) {
        compute_cutoff_list(kmc_comm.comm_pairwise, kmc_comm.counts_pairwise.data(), kmc_comm.displs_pairwise.data(), gpubuf);
        std::cout << "max num cutoff " << gpubuf.N_cutoff_ << std::endl;
        initialize_sparsity_K(gpubuf, p.pbc, p.nn_dist, NL, kmc_comm);
    }
    copytoConstMemory(gpubuf, std::vector<double>(E_gen, E_gen + 5), std::vector<double>(E_rec, E_rec + 5),
                      std::vector<double>(E_Vdiff, E_Vdiff + 5), std::vector<double>(E_Odiff, E_Odiff + 5));

    // ---- bias loop (src/kmc_main.cpp:257-575) -----------------------------------------------------------------------
    std::vector<double> V_switch(std::max(1, p.n_V_switch)), t_switch(std::max(1, p.n_t_switch));
    kmcb200_parse_parameter_vector(argv[1], 0, (int)V_switch.size(), V_switch.data());
    kmcb200_parse_parameter_vector(argv[1], 1, (int)t_switch.size(), t_switch.data());
    auto tcode_start = std::chrono::steady_clock::now();
    for (size_t vt = 0; vt < V_switch.size() && vt < t_switch.size(); vt++) {
        const double Vd = V_switch[vt], t = t_switch[vt];
        outputBuffer << "--------------------------------\n" << "Applied Voltage = " << Vd << " V\n"
                     << "--------------------------------\n";
        const std::string folder = "Results_" + std::to_string(Vd);
        mkdir(folder.c_str(), S_IRWXU | S_IRWXG | S_IROTH | S_IXOTH);
        outputBuffer << "Created folder: " << folder << '\n';
        write_snapshot(device, folder, "snapshot_init.xyz");
        double kmc_time = 0.0;
        long kmc_step_count = 0;
        gpubuf.h2d((int *)gpubuf.site_element, device.site_element.data(), device.N);  // sync_HostToGPU
        gpubuf.h2d(gpubuf.site_charge, device.site_charge.data(), device.N);
        while (kmc_time < t && (max_steps < 0 || kmc_step_count < max_steps)) {
            auto t0 = std::chrono::steady_clock::now();
            if (p.solve_potential) {
                update_charge_gpu(gpubuf.site_element, gpubuf.site_charge, gpubuf.neigh_idx, gpubuf.N_, gpubuf.nn_, metals,
                                  kmc_comm.counts_events.data(), kmc_comm.displs_events.data(), kmc_comm.comm_events, ctx);
                background_potential_gpu_sparse(nullptr, nullptr, gpubuf, device.N, NL, NL, Vd, p.pbc, p.high_G, p.low_G,
                                                p.nn_dist, p.num_metals, (int)kmc_step_count);
                poisson_gridless_gpu(ctx, p.num_atoms_contact, p.pbc, gpubuf.N_, gpubuf.sigma_h, gpubuf.k_h, gpubuf.site_x,
                                     gpubuf.site_y, gpubuf.site_z, gpubuf.site_element, gpubuf.site_charge,
                                     gpubuf.site_potential_charge, kmc_comm.rank_pairwise, kmc_comm.size_pairwise,
                                     kmc_comm.counts_pairwise.data(), kmc_comm.displs_pairwise.data());
                sum_and_gather_potential(gpubuf, NL, kmc_comm);
            }
            if (p.perturb_structure) {
                double event_time = execute_kmc_step_mpi(kmc_comm.comm_events, gpubuf, device.N, kmc_comm.counts_events.data(),
                                                         kmc_comm.displs_events.data(), 52, gpubuf.neigh_idx, gpubuf.site_layer,
                                                         p.pbc, gpubuf.T_bg_h, gpubuf.freq_h, gpubuf.sigma_h, gpubuf.k_h,
                                                         gpubuf.site_x, gpubuf.site_y, gpubuf.site_z, gpubuf.site_potential_charge,
                                                         gpubuf.site_element, gpubuf.site_charge, random_generator);
                kmc_time += event_time;
                std::cout << "Number of KMC events: " << gpubuf.last_n_events << "\n";
            } else if (kmc_step_count > 0) {
                kmc_time = t;
            }
            double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            outputBuffer << "KMC time is: " << kmc_time << "\n";
            if (!(kmc_step_count % std::max(1, p.output_freq))) { outputFile << outputBuffer.str(); outputBuffer.str(std::string()); }
            kmc_step_count++;
            outputBuffer << "Z - calculation time - KMC superstep [s]: " << dt << "\n" << "PCG iterations: "
                         << gpubuf.last_cg_iterations << "\n--------------------------------------\n";
        }
        // sync_GPUToHost + final snapshot (src/kmc_main.cpp:556-565)
        gpubuf.d2h(device.site_element.data(), (const int *)gpubuf.site_element, device.N);
        gpubuf.d2h(device.site_charge.data(), gpubuf.site_charge, device.N);
        gpubuf.d2h(device.site_potential_charge.data(), gpubuf.site_potential_charge, device.N);
        KMCB200_CHECK(kmcb200_synchronize(ctx));
        std::cout << "KMC step count: " << kmc_step_count << "\n";
        write_snapshot(device, folder, "snapshot_" + std::to_string(kmc_step_count) + ".xyz");
    }
    std::cout << "Total code execution time: "
              << std::chrono::duration<double>(std::chrono::steady_clock::now() - tcode_start).count() << " s\n";
    outputFile << outputBuffer.str();
    outputFile.close();
    return 0;
}
