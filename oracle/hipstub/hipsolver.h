#pragma once
typedef void *hipsolverHandle_t;
typedef void *hipsolverDnHandle_t;
typedef int hipsolverStatus_t;
#define HIPSOLVER_STATUS_SUCCESS 0
