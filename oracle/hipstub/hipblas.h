#pragma once
typedef void *hipblasHandle_t;
typedef int hipblasStatus_t;
#define HIPBLAS_STATUS_SUCCESS 0
