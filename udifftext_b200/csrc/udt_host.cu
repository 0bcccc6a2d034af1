// udt_host.cu — error reporting, device checks and TMA tensor-map encoding for libudt_b200.so.
#include "udt_host.h"

#include <mutex>
#include <stdlib.h>
#include <string.h>

namespace udt_host {

static thread_local char g_err[512] = "";

char* error_buffer() { return g_err; }

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(UDT_ERR_LAUNCH, "%s: %s", what, cudaGetErrorString(e));
  return UDT_OK;
}

static int g_arch[64];
static int g_sms[64];
static bool g_arch_known[64];

static int query_device(int* dev_out) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return fail(UDT_ERR_ARCH, "cudaGetDevice: %s (no CUDA device; this library has no CPU path)",
                                    cudaGetErrorString(e));
  if (dev < 0 || dev >= 64) return fail(UDT_ERR_ARCH, "device index %d out of range", dev);
  if (!g_arch_known[dev]) {
    int major = 0, minor = 0, sms = 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    g_arch[dev] = major * 10 + minor;
    g_sms[dev] = sms;
    g_arch_known[dev] = true;
  }
  *dev_out = dev;
  return UDT_OK;
}

int require_sm100() {
  int dev;
  int rc = query_device(&dev);
  if (rc != UDT_OK) return rc;
  if (g_arch[dev] != 100)
    return fail(UDT_ERR_ARCH, "device %d is sm_%d; libudt_b200 is built for sm_100a only (no fallback)", dev,
                g_arch[dev]);
  return UDT_OK;
}

int num_sms() {
  int dev;
  if (query_device(&dev) != UDT_OK) return -1;
  return g_sms[dev];
}

int arch() {
  int dev;
  int rc = query_device(&dev);
  if (rc != UDT_OK) return rc;
  return g_arch[dev];
}

int pdl_attr(cudaLaunchAttribute* attr) {
  static const int enabled = tune_int("UDT_PDL", 1);
  if (!enabled) return 0;
  attr->id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr->val.programmaticStreamSerializationAllowed = 1;
  return 1;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

static int encode(CUtensorMap* m, const void* ptr, uint32_t rank, const cuuint64_t* dims, const cuuint64_t* strides,
                  const cuuint32_t* box, const cuuint32_t* estr_in = nullptr,
                  CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
  EncodeTiledFn fn = get_encode();
  if (!fn) return fail(UDT_ERR_DRIVER, "cuTensorMapEncodeTiled entry point not available");
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0) return fail(UDT_ERR_ALIGN, "TMA base pointer not 16-byte aligned");
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  if (estr_in != nullptr)
    for (uint32_t i = 0; i < rank; ++i) estr[i] = estr_in[i];
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, rank, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(UDT_ERR_DRIVER, "cuTensorMapEncodeTiled failed (CUresult %d; rank %u dims %llu,%llu box %u,%u)", (int)r,
                rank, (unsigned long long)dims[0], (unsigned long long)dims[1], box[0], box[1]);
  return UDT_OK;
}

int make_tmap_2d(CUtensorMap* m, const void* ptr, uint64_t cols, uint64_t rows, uint64_t ld_elems, uint32_t box0,
                 uint32_t box1) {
  if ((ld_elems * 2) % 16 != 0) return fail(UDT_ERR_ALIGN, "row pitch %llu elements is not a multiple of 8", (unsigned long long)ld_elems);
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld_elems * 2};
  cuuint32_t box[2] = {box0, box1};
  return encode(m, ptr, 2, dims, strides, box);
}

int make_tmap_nhwc(CUtensorMap* m, const void* ptr, uint64_t C, uint64_t W, uint64_t H, uint64_t N, uint64_t ld,
                   uint32_t bw, uint32_t bh, uint32_t bn, uint32_t pix_stride) {
  if ((ld * 2) % 16 != 0) return fail(UDT_ERR_ALIGN, "channel pitch %llu elements is not a multiple of 8", (unsigned long long)ld);
  cuuint64_t dims[4] = {C, W, H, N};
  cuuint64_t strides[3] = {ld * 2, ld * 2 * W, ld * 2 * W * H};
  // a box of bw x bh *loaded* pixels spans bw*stride x bh*stride tensor elements (TMA traversal stride)
  cuuint32_t box[4] = {64, bw * pix_stride, bh * pix_stride, bn};
  cuuint32_t estr[4] = {1, pix_stride, pix_stride, 1};
  return encode(m, ptr, 4, dims, strides, box, estr);
}

// epilogue tile map: {32 channels, bw, bh, bn} boxes of an NHWC fp16 tensor with the 64B swizzle (rows of 64 B)
int make_tmap_nhwc_c32(CUtensorMap* m, const void* ptr, uint64_t C, uint64_t W, uint64_t H, uint64_t N, uint64_t ld,
                       uint32_t bw, uint32_t bh, uint32_t bn) {
  if ((ld * 2) % 16 != 0) return fail(UDT_ERR_ALIGN, "channel pitch %llu elements is not a multiple of 8", (unsigned long long)ld);
  cuuint64_t dims[4] = {C, W, H, N};
  cuuint64_t strides[3] = {ld * 2, ld * 2 * W, ld * 2 * W * H};
  cuuint32_t box[4] = {32, bw, bh, bn};
  return encode(m, ptr, 4, dims, strides, box, nullptr, CU_TENSOR_MAP_SWIZZLE_64B);
}

int make_tmap_nhwc_c32_strided(CUtensorMap* m, const void* ptr, uint64_t C, uint64_t W, uint64_t H, uint64_t N, uint64_t sw,
                               uint64_t sh, uint64_t sn, uint32_t bw, uint32_t bh, uint32_t bn) {
  if ((sw * 2) % 16 != 0 || (sh * 2) % 16 != 0 || (sn * 2) % 16 != 0)
    return fail(UDT_ERR_ALIGN, "strided view: pixel strides must be multiples of 8 elements");
  cuuint64_t dims[4] = {C, W, H, N};
  cuuint64_t strides[3] = {sw * 2, sh * 2, sn * 2};
  cuuint32_t box[4] = {32, bw, bh, bn};
  return encode(m, ptr, 4, dims, strides, box, nullptr, CU_TENSOR_MAP_SWIZZLE_64B);
}

}  // namespace udt_host

extern "C" {
int udt_version(void) { return 5; }
int udt_arch(void) { return udt_host::arch(); }
const char* udt_last_error(void) { return udt_host::error_buffer(); }
int udt_num_sms(void) { return udt_host::num_sms(); }
int udt_sizeof_igemm_desc(void) { return static_cast<int>(sizeof(udt_igemm_desc)); }
}
