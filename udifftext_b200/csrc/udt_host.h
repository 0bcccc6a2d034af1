// udt_host.h — host-side helpers shared by the C-ABI translation units (error reporting, device query,
// TMA tensor-map encoding through the driver entry point so the library does not link libcuda).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <string.h>

#include "../../include/udt_api.h"

namespace udt_host {

char* error_buffer();  // thread-local, 512 bytes
int fail(int code, const char* fmt, ...);
int check_launch(const char* what);   // cudaGetLastError -> UDT_ERR_LAUNCH
int require_sm100();                  // UDT_OK or UDT_ERR_ARCH (cached per device)
int num_sms();
int arch();                           // compute capability * 10 of the current device, or <0
// Experiment switches.  A PRODUCTION build of this library reads no environment variable and has no hidden state: every
// switch is the compile-time default below.  A TUNING build (`UDT_TRACE=1 python -m udifftext_b200.build`, which defines
// UDT_TUNING) reads the named variable once, so that A/B measurements (scripts/) need no rebuild per variant.
#ifdef UDT_TUNING
#include <stdlib.h>
inline int tune_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e != nullptr ? atoi(e) : dflt;
}
#else
inline constexpr int tune_int(const char*, int dflt) { return dflt; }
#endif

// Programmatic dependent launch (every kernel of this library executes griddepcontrol.wait before it touches
// global memory): fills one launch attribute and returns 1, or returns 0 when disabled (tuning builds: UDT_PDL=0).
int pdl_attr(cudaLaunchAttribute* attr);

// kernel<<<grid, block, smem, stream>>>(args...) with the programmatic-dependent-launch attribute
template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  cfg.numAttrs = pdl_attr(attr);
  cfg.attrs = attr;
  (void)cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);   // errors surface through check_launch()
}

// 2-D fp16 tensor map, dim0 = contiguous (cols), box = {box0, box1}, 128B swizzle.
int make_tmap_2d(CUtensorMap* m, const void* ptr, uint64_t cols, uint64_t rows, uint64_t ld_elems, uint32_t box0,
                 uint32_t box1);
// 4-D fp16 NHWC tensor map: dims (C, W, H, N), channel pitch `ld` elements, box = {64, bw, bh, bn} loaded pixels
// taken every `pix_stride`-th pixel along W and H (TMA element strides), 128B swizzle.
int make_tmap_nhwc(CUtensorMap* m, const void* ptr, uint64_t C, uint64_t W, uint64_t H, uint64_t N, uint64_t ld,
                   uint32_t bw, uint32_t bh, uint32_t bn, uint32_t pix_stride);

// epilogue tile map: {32 channels, bw, bh, bn} boxes, 64B swizzle (TMA store of finished tiles / residual prefetch)
int make_tmap_nhwc_c32(CUtensorMap* m, const void* ptr, uint64_t C, uint64_t W, uint64_t H, uint64_t N, uint64_t ld,
                       uint32_t bw, uint32_t bh, uint32_t bn);

// same boxes over a strided pixel view: pixel (w, h, n) at element offset w*sw + h*sh + n*sn (all multiples of 8)
int make_tmap_nhwc_c32_strided(CUtensorMap* m, const void* ptr, uint64_t C, uint64_t W, uint64_t H, uint64_t N, uint64_t sw,
                               uint64_t sh, uint64_t sn, uint32_t bw, uint32_t bh, uint32_t bn);

}  // namespace udt_host
