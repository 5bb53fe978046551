// Status codes shared by every C-ABI entry point (mirrored in include/groma_b200.h).
#pragma once
#include <stdint.h>

#define GROMA_OK 0
#define GROMA_ERR_ARG 1         // null pointer / bad shape
#define GROMA_ERR_ALIGN 2       // pointer or leading dimension not 16-byte aligned
#define GROMA_ERR_CUDA 3        // launch failed (cudaGetLastError)
#define GROMA_ERR_DRIVER 4      // driver entry point unavailable
#define GROMA_ERR_TMA_ENCODE 5  // cuTensorMapEncodeTiled rejected the descriptor
#define GROMA_ERR_UNSUPPORTED 6 // shape outside what the kernels are compiled for

#ifdef __cplusplus
#define GROMA_API extern "C" __attribute__((visibility("default")))
#else
#define GROMA_API __attribute__((visibility("default")))
#endif

#define GROMA_LAUNCH_CHECK() (cudaGetLastError() == cudaSuccess ? GROMA_OK : GROMA_ERR_CUDA)
