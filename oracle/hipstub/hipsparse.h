#pragma once
typedef void *hipsparseHandle_t;
