// kmc_main.cpp -- host driver with the call order of the reference's main() (src/kmc_main.cpp:56-603), written against
// include/gpu_solvers_b200.hpp: every solver call below has the reference's name and the reference's argument list
// (compare src/kmc_main.cpp:187-239 and :328-540).  Reads parameters.txt and the xyz structure files unchanged; writes
// output<size>_<rank>.txt with the reference's "KMC time is:" lines and Results_<V>/snapshot_*.xyz in the reference's
// snapshot format (src/Device.cpp:214-232), so a run can be diffed against structures/5nm_device/expected_output.
// The Kirchhoff / current chain (dead in the reference's shipped main, src/KMC_comm.h:243) runs when solve_current is set
// in parameters.txt AND KMCB200_ENABLE_CURRENT=1: the macroscopic current is then logged every superstep.
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

// ---- host model: the members of the reference's KMCParameters / Device / KMCProcess that main() uses -----------------
class KMCParameters {  // src/input_parser.h
public:
    kmcb200_params raw;
    bool solve_potential, solve_current, solve_heating_local, solve_heating_global, perturb_structure, pristine, restart;
    int pbc, num_atoms_first_layer, num_atoms_contact, num_layers_contact, output_freq;
    unsigned rnd_seed;
    double high_G, low_G, nn_dist, m_e, V0, q = 1.60217663e-19, sigma, k, freq, background_temp, initial_vacancy_concentration;
    std::vector<ELEMENT> metals;
    std::vector<double> V_switch, t_switch, lattice;
    std::string restart_xyz_file, atom_xyz_file, interstitial_xyz_file;
    explicit KMCParameters(const std::string &file) {
        KMCB200_CHECK(kmcb200_parse_parameters(file.c_str(), &raw));
        solve_potential = raw.solve_potential; solve_current = raw.solve_current;
        solve_heating_local = raw.solve_heating_local; solve_heating_global = raw.solve_heating_global;
        perturb_structure = raw.perturb_structure; pristine = raw.pristine; restart = raw.restart;
        pbc = raw.pbc; num_atoms_first_layer = raw.num_atoms_first_layer; num_atoms_contact = raw.num_atoms_contact;
        num_layers_contact = raw.num_layers_contact; output_freq = std::max(1, raw.output_freq); rnd_seed = raw.rnd_seed;
        high_G = raw.high_G; low_G = raw.low_G; nn_dist = raw.nn_dist; V0 = raw.V0;
        m_e = raw.m_r * 9.11e-31;  // src/input_parser.cpp:397, src/input_parser.h:99
        sigma = raw.sigma; k = raw.k; freq = raw.freq; background_temp = raw.background_temp;
        initial_vacancy_concentration = raw.initial_vacancy_concentration;
        for (int i = 0; i < raw.num_metals; ++i) metals.push_back((ELEMENT)raw.metals[i]);
        V_switch.resize(std::max(1, raw.n_V_switch)); t_switch.resize(std::max(1, raw.n_t_switch));
        kmcb200_parse_parameter_vector(file.c_str(), 0, (int)V_switch.size(), V_switch.data());
        kmcb200_parse_parameter_vector(file.c_str(), 1, (int)t_switch.size(), t_switch.data());
        lattice.assign(raw.lattice, raw.lattice + 3);
        restart_xyz_file = raw.restart_xyz_file; atom_xyz_file = raw.atom_xyz_file;
        interstitial_xyz_file = raw.interstitial_xyz_file;
    }
};

class Device {  // src/Device.h
public:
    int N = 0, N_atom = 0, max_num_neighbors = 52, pbc = 0;  // max_num_neighbors: src/Device.cpp:59
    double nn_dist = 3.5, sigma = 0, k = 0, imacro = 0, T_bg = 300;
    std::vector<ELEMENT> site_element;
    std::vector<int> site_charge;
    std::vector<double> site_x, site_y, site_z, lattice, site_power, site_CB_edge, site_potential_boundary,
        site_potential_charge, site_temperature;

    Device(const std::vector<std::string> &xyz_files, KMCParameters &p) {  // src/Device.cpp:17-74
        for (auto &path : xyz_files) {
            int n = kmcb200_xyz_count(path.c_str());
            if (n < 0) { std::fprintf(stderr, "%s\n", kmcb200_last_error()); std::exit(1); }
            size_t o = site_element.size();
            site_element.resize(o + n); site_x.resize(o + n); site_y.resize(o + n); site_z.resize(o + n);
            if (kmcb200_read_xyz(path.c_str(), n, (int *)site_element.data() + o, site_x.data() + o, site_y.data() + o,
                                 site_z.data() + o) < 0) {
                std::fprintf(stderr, "%s\n", kmcb200_last_error());
                std::exit(1);
            }
        }
        N = (int)site_element.size();
        pbc = p.pbc; nn_dist = p.nn_dist; sigma = p.sigma; k = p.k; T_bg = p.background_temp; lattice = p.lattice;
        site_charge.assign(N, 0);
        for (auto *v : {&site_power, &site_CB_edge, &site_potential_boundary, &site_potential_charge}) v->assign(N, 0.0);
        site_temperature.assign(N, T_bg);
        updateAtomLists();
        std::cout << "Loaded " << N << " sites into device\n";
    }
    void updateAtomLists() {  // src/Device.cpp:116-143
        N_atom = 0;
        for (int i = 0; i < N; i++)
            if (site_element[i] != DEFECT && site_element[i] != OXYGEN_DEFECT) N_atom++;
    }
    void makeSubstoichiometric(double vacancy_concentration, unsigned rnd_seed) {  // src/Device.cpp:180-211
        int nv = kmcb200_make_substoichiometric(N, (int *)site_element.data(), vacancy_concentration, rnd_seed);
        std::cout << nv << " oxygen atoms will be converted to vacancies" << std::endl;
    }
    void writeSnapshot(std::string filename, std::string foldername) {  // src/Device.cpp:214-232
        std::ofstream fout(("./" + foldername + "/" + filename).c_str());
        fout << N << "\n\n";
        for (int i = 0; i < N; i++)
            fout << kElementNames[site_element[i]] << "   " << site_x[i] << "   " << site_y[i] << "   " << site_z[i] << "   "
                 << site_potential_charge[i] << "   " << site_power[i] << "\n";
    }
    void setLaplacePotential(hipblasHandle_t handle_cublas, hipsolverHandle_t handle_cusolver, GPUBuffers &gpubuf,
                             KMCParameters &p, double Vd) {  // src/potential_solver.cpp:4-18
        gpubuf.sync_HostToGPU(*this);
        update_CB_edge_gpu_sparse(handle_cublas, handle_cusolver, gpubuf, N, p.num_atoms_first_layer, p.num_atoms_first_layer,
                                  Vd, pbc, p.high_G, p.low_G, nn_dist, (int)p.metals.size());
        gpubuf.sync_GPUToHost(*this);
    }
};
void GPUBuffers::sync_HostToGPU(Device &device) { kmcb200_sync_host_to_gpu(*this, device); }  // src/gpu_buffers.cpp:10-34
void GPUBuffers::sync_GPUToHost(Device &device) { kmcb200_sync_gpu_to_host(*this, device); }  // src/gpu_buffers.cpp:36-55

class KMCProcess {  // src/KMCProcess.h, src/KMCProcess.cpp:17-50
public:
    std::vector<Layer> layers;
    std::vector<int> site_layer;
    double freq;
    RandomNumberGenerator random_generator;
    KMCProcess(Device &device, double _freq) : freq(_freq) {
        random_generator.setSeed(1);  // rnd_seed_kmc, src/structure_input.h:8
        double E_gen[5], E_rec[5], E_Vdiff[5], E_Odiff[5], sx[5], ex[5];
        const int nl = kmcb200_num_layers();
        kmcb200_layer_table(E_gen, E_rec, E_Vdiff, E_Odiff, sx, ex);
        layers.resize(nl);
        for (int l = 0; l < nl; ++l) layers[l].init_layer("", E_gen[l], E_rec[l], E_Vdiff[l], E_Odiff[l], sx[l], ex[l]);
        site_layer.resize(device.N);
        KMCB200_CHECK(kmcb200_assign_layers(device.N, device.site_x.data(), site_layer.data()));
    }
};

int main(int argc, char **argv) {
    if (argc < 2) { std::fprintf(stderr, "usage: %s parameters.txt [max_supersteps]\n", argv[0]); return 2; }
    const long max_steps = argc > 2 ? std::atol(argv[2]) : -1;
    MPI_Init(&argc, &argv);
    int rank_global = 0, size_global = 1;
    MPI_Comm_rank(MPI_COMM_WORLD, &rank_global);
    MPI_Comm_size(MPI_COMM_WORLD, &size_global);

    KMCParameters p(argv[1]);
    std::string dir(argv[1]);
    dir = dir.find_last_of('/') == std::string::npos ? "." : dir.substr(0, dir.find_last_of('/'));
    std::ofstream outputFile("output" + std::to_string(size_global) + "_" + std::to_string(rank_global) + ".txt");
    std::ostringstream outputBuffer;

    // ---- Device (src/kmc_main.cpp:127-155) ------------------------------------------------------------------
    std::vector<std::string> xyz_files;
    if (p.restart) {
        outputBuffer << "Restarting from " << p.restart_xyz_file << "\n";
        xyz_files.push_back(dir + "/" + p.restart_xyz_file);
    } else {
        xyz_files.push_back(dir + "/" + p.atom_xyz_file);
        xyz_files.push_back(dir + "/" + p.interstitial_xyz_file);
    }
    Device device(xyz_files, p);
    if (p.pristine) device.makeSubstoichiometric(p.initial_vacancy_concentration, p.rnd_seed);

    // ---- communicators, KMC process, GPU buffers (src/kmc_main.cpp:161-191) ---------------------------------
    bool split = false;
    int ratio[2] = {8, 24};
    KMC_comm kmc_comm(MPI_COMM_WORLD, device.N - 2 * p.num_atoms_first_layer, device.N_atom + 1, device.N, device.N, split, ratio);
    KMCProcess sim(device, p.freq);
    GPUBuffers gpubuf(sim.layers, sim.site_layer, sim.freq, device.N, device.N_atom, device.site_element, device.site_x,
                      device.site_y, device.site_z, device.max_num_neighbors, device.sigma, device.k, device.lattice, p.metals,
                      p.metals.size(), MPI_COMM_WORLD, p.num_atoms_first_layer);

    // ---- neighbour lists + K sparsity (src/kmc_main.cpp:197-218) --------------------------------------------
    if (kmc_comm.comm_events != MPI_COMM_NULL)
        compute_neighbor_list(kmc_comm.comm_events, kmc_comm.counts_events, kmc_comm.displs_events, device, gpubuf, p);
    if (p.solve_potential && kmc_comm.comm_pairwise != MPI_COMM_NULL) {
        compute_cutoff_list(kmc_comm.comm_pairwise, kmc_comm.counts_pairwise, kmc_comm.displs_pairwise, device, gpubuf, p);
        std::cout << "max num cutoff " << gpubuf.N_cutoff_ << std::endl;
    }
    gpubuf.sync_HostToGPU(device);
    if (p.solve_potential && kmc_comm.comm_K != MPI_COMM_NULL)
        initialize_sparsity_K(gpubuf, p.pbc, p.nn_dist, p.num_atoms_first_layer, kmc_comm);
    if (p.solve_current && kmc_comm.comm_T != MPI_COMM_NULL) initialize_sparsity_CB(gpubuf, p.pbc, p.nn_dist, p.num_atoms_first_layer);
    std::vector<double> E_gen_host, E_rec_host, E_Vdiff_host, E_Odiff_host;
    for (auto l : sim.layers) {
        E_gen_host.push_back(l.E_gen_0); E_rec_host.push_back(l.E_rec_1);
        E_Vdiff_host.push_back(l.E_diff_2); E_Odiff_host.push_back(l.E_diff_3);
    }
    copytoConstMemory(E_gen_host, E_rec_host, E_Vdiff_host, E_Odiff_host);
    hipblasHandle_t handle;
    hipblasCreate(&handle);
    hipsolverHandle_t handle_cusolver;
    hipsolverCreate(&handle_cusolver);

    // ---- bias loop (src/kmc_main.cpp:257-575) ---------------------------------------------------------------
    auto tcode_start = std::chrono::steady_clock::now();
    for (size_t vt_counter = 0; vt_counter < p.V_switch.size() && vt_counter < p.t_switch.size(); vt_counter++) {
        const double Vd = p.V_switch[vt_counter], t = p.t_switch[vt_counter];
        outputBuffer << "--------------------------------\n" << "Applied Voltage = " << Vd << " V\n"
                     << "--------------------------------\n";
        if (p.solve_current && kmc_comm.comm_T != MPI_COMM_NULL) {  // src/kmc_main.cpp:270-275
            device.setLaplacePotential(handle, handle_cusolver, gpubuf, p, Vd);
            initialize_sparsity_T(gpubuf, p.pbc, p.nn_dist, p.num_atoms_first_layer, p.num_atoms_first_layer, p.num_layers_contact, kmc_comm);
            std::cout << "Initialized sparsity of T\n";
        }
        const std::string folder_name = "Results_" + std::to_string(Vd);
        if (rank_global == 0) {
            mkdir(folder_name.c_str(), S_IRWXU | S_IRWXG | S_IROTH | S_IXOTH);
            outputBuffer << "Created folder: " << folder_name << '\n';
            device.writeSnapshot("snapshot_init.xyz", folder_name);
        }
        // constants from parameterization (src/kmc_main.cpp:294-302)
        double loop_G = p.high_G * 10000000, high_G = p.high_G * 100000, low_G = p.low_G;
        double scale = 1e-5, G0 = 2 * 3.8612e-5 * scale, tol = p.q * 0.01, alpha = 1;
        int num_source_inj = p.num_atoms_first_layer, num_ground_ext = p.num_atoms_first_layer;
        double kmc_time = 0.0;
        int kmc_step_count = 0;
        gpubuf.sync_HostToGPU(device);
        while (kmc_time < t && (max_steps < 0 || kmc_step_count < max_steps)) {
            MPI_Barrier(kmc_comm.comm_events);
            double t_superstep_start = MPI_Wtime();
            if (p.solve_potential) {
                if (kmc_comm.comm_events != MPI_COMM_NULL)
                    update_charge_gpu(gpubuf.site_element, gpubuf.site_charge, gpubuf.neigh_idx, gpubuf.N_, gpubuf.nn_,
                                      gpubuf.metal_types, gpubuf.num_metal_types_, kmc_comm.counts_events,
                                      kmc_comm.displs_events, kmc_comm.comm_events);
                if (kmc_comm.comm_K != MPI_COMM_NULL)
                    background_potential_gpu_sparse(handle, handle_cusolver, gpubuf, device.N, p.num_atoms_first_layer,
                                                    p.num_atoms_first_layer, Vd, p.pbc, p.high_G, p.low_G, device.nn_dist,
                                                    p.metals.size(), kmc_step_count);
                if (kmc_comm.comm_pairwise != MPI_COMM_NULL)
                    poisson_gridless_gpu(p.num_atoms_contact, p.pbc, gpubuf.N_, gpubuf.lattice, gpubuf.sigma, gpubuf.k,
                                         gpubuf.site_x, gpubuf.site_y, gpubuf.site_z, gpubuf.site_charge,
                                         gpubuf.site_potential_charge, kmc_comm.rank_pairwise, kmc_comm.size_pairwise,
                                         kmc_comm.counts_pairwise, kmc_comm.displs_pairwise, gpubuf.cutoff_window,
                                         gpubuf.cutoff_idx, gpubuf.N_cutoff_);
            }
            if (p.solve_current && kmc_comm.comm_T != MPI_COMM_NULL) {
                update_power_gpu_sparse_dist(handle, handle_cusolver, gpubuf, num_source_inj, num_ground_ext, p.num_layers_contact,
                                             Vd, high_G, low_G, loop_G, G0, tol, device.nn_dist, p.m_e, p.V0, p.metals.size(),
                                             &device.imacro, p.solve_heating_local, p.solve_heating_global, alpha);
                outputBuffer << "I_macro: " << device.imacro * (1e6) << "\n";  // src/current_solver_gpu.cu:2055
            }
            if (p.solve_potential) sum_and_gather_potential(gpubuf, p.num_atoms_first_layer, kmc_comm);
            if (p.perturb_structure) {
                if (kmc_comm.comm_events != MPI_COMM_NULL) {
                    double event_time = execute_kmc_step_mpi(
                        kmc_comm.comm_events, device.N, kmc_comm.counts_events, kmc_comm.displs_events, device.max_num_neighbors,
                        gpubuf.neigh_idx, gpubuf.site_layer, gpubuf.lattice, device.pbc, gpubuf.T_bg, gpubuf.freq, gpubuf.sigma,
                        gpubuf.k, gpubuf.site_x, gpubuf.site_y, gpubuf.site_z, gpubuf.site_potential_charge,
                        gpubuf.site_temperature, gpubuf.site_element, gpubuf.site_charge, sim.random_generator);
                    kmc_time += event_time;
                    std::cout << "Number of KMC events: " << kmcb200::rt().last_n_events << "\n";
                }
            } else if (kmc_step_count > 0) {
                kmc_time = t;
            }
            double t_superstep_end = MPI_Wtime();
            outputBuffer << "KMC time is: " << kmc_time << "\n";
            if (!(kmc_step_count % p.output_freq)) { outputFile << outputBuffer.str(); outputBuffer.str(std::string()); }
            kmc_step_count++;
            outputBuffer << "Z - calculation time - KMC superstep [s]: " << t_superstep_end - t_superstep_start << "\n"
                         << "PCG iterations: " << kmcb200::rt().last_cg_iterations << "\n--------------------------------------\n";
        }
        hipDeviceSynchronize();
        gpubuf.sync_GPUToHost(device);  // src/kmc_main.cpp:556-565
        std::cout << "KMC step count: " << kmc_step_count << "\n";
        if (rank_global == 0) device.writeSnapshot("snapshot_" + std::to_string(kmc_step_count) + ".xyz", folder_name);
    }
    std::cout << "Total code execution time: "
              << std::chrono::duration<double>(std::chrono::steady_clock::now() - tcode_start).count() << " s\n";
    outputFile << outputBuffer.str();
    outputFile.close();
    MPI_Finalize();
    return 0;
}
