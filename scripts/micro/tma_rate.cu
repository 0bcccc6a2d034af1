// tma_rate.cu — micro-benchmark (tuning aid, not product): per-SM TMA load / store throughput as a function of the
// box row length (64 B vs 128 B), with all SMs or a subset active, L2-resident data.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I udifftext_b200/csrc scripts/micro/tma_rate.cu udifftext_b200/csrc/udt_host.cu -o gpurun_out/tma_rate
#include "udt_common.cuh"
#include "udt_host.h"
#include <vector>

using namespace udt;

struct P {
  CUtensorMap map;
  int box_bytes, rows_per_box, col_chunks, row_blocks, nloads, stages, store, nprod;
  unsigned long long* out;
};

__global__ void __launch_bounds__(128, 1) k_rate(const __grid_constant__ P p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t basea = (raw + 1023u) & ~1023u;
  uint8_t* base = smem_raw + (basea - raw);
  uint64_t* bar = reinterpret_cast<uint64_t*>(base);
  uint8_t* buf = base + 1024;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&p.map);
    for (int s = 0; s < p.stages * p.nprod; ++s) mbar_init(&bar[s], 1);
    fence_mbar_init();
  }
  __syncthreads();
  const int pw = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0 && pw < p.nprod) {
    bar += pw * p.stages;
    buf += pw * p.stages * p.box_bytes;
    const long long t0 = clock64();
    const int row_base = blockIdx.x * p.row_blocks * p.rows_per_box;
    (void)pw;
    if (!p.store) {
      const int box_cols = p.box_bytes / p.rows_per_box / 2;
      int s = 0, c = 0, rb = 0;
      uint32_t ph = 1;   // parity to wait for = (round - 1) & 1; round 0 never waits
      bool first_round = true;
      for (int i = 0; i < p.nloads; ++i) {
        if (!first_round) mbar_wait(&bar[s], ph);
        mbar_expect_tx(&bar[s], p.box_bytes);
        tma_load_2d(&p.map, &bar[s], buf + s * p.box_bytes, c * box_cols, row_base + rb * p.rows_per_box);
        if (++c == p.col_chunks) { c = 0; if (++rb == p.row_blocks) rb = 0; }
        if (++s == p.stages) { s = 0; if (first_round) { first_round = false; ph = 0; } else ph ^= 1u; }
      }
      for (int k = 0; k < p.stages; ++k) {   // drain
        if (!first_round) mbar_wait(&bar[s], ph);
        if (++s == p.stages) { s = 0; if (first_round) { first_round = false; ph = 0; } else ph ^= 1u; }
      }
    } else {
      const int box_cols = p.box_bytes / p.rows_per_box / 2;
      int s = 0, c = 0, rb = 0;
      for (int i = 0; i < p.nloads; ++i) {
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                         reinterpret_cast<uint64_t>(&p.map)),
                     "r"(smem_u32(buf + s * p.box_bytes)), "r"(c * box_cols), "r"(row_base + rb * p.rows_per_box)
                     : "memory");
        tma_store_commit();
        asm volatile("cp.async.bulk.wait_group.read 6;" ::: "memory");
        if (++c == p.col_chunks) { c = 0; if (++rb == p.row_blocks) rb = 0; }
        if (++s == p.stages) s = 0;
      }
      tma_store_wait_all<0>();
    }
    const long long t1 = clock64();
    if (pw == 0) p.out[blockIdx.x] = static_cast<unsigned long long>(t1 - t0);
  }
}

static int run(const char* label, int grid, int cols, int box_cols, int rows_per_box, int store, CUtensorMapSwizzle swz, int stages = 8, int nprod = 1) {
  using namespace udt_host;
  const int col_chunks = cols / box_cols;
  const int row_blocks = 4;
  const int rows_total = 148 * row_blocks * rows_per_box;
  __half* d;
  cudaMalloc(&d, static_cast<size_t>(rows_total) * cols * 2);
  cudaMemset(d, 0, static_cast<size_t>(rows_total) * cols * 2);
  P p;
  memset(&p, 0, sizeof(p));
  // encode through the driver entry point directly (need the swizzle choice)
  void* fnp = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
  typedef CUresult (*Fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                         const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                         CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows_total};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)rows_per_box};
  cuuint32_t es[2] = {1, 1};
  CUresult r = ((Fn)fnp)(&p.map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, d, dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("%s: encode failed %d\n", label, (int)r); return 1; }
  p.box_bytes = box_cols * 2 * rows_per_box;
  p.rows_per_box = rows_per_box;
  p.col_chunks = col_chunks;
  p.row_blocks = row_blocks;
  p.nloads = 400;
  p.stages = stages;
  p.nprod = nprod;
  p.store = store;
  cudaMalloc(&p.out, 148 * 8);
  const int smem = 2048 + p.stages * p.box_bytes * nprod;
  cudaFuncSetAttribute(k_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  for (int it = 0; it < 3; ++it) k_rate<<<grid, 128, smem>>>(p);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: %s\n", label, cudaGetErrorString(e)); return 1; }
  std::vector<unsigned long long> h(148);
  cudaMemcpy(h.data(), p.out, grid * 8, cudaMemcpyDeviceToHost);
  double avg = 0;
  unsigned long long mx = 0;
  for (int i = 0; i < grid; ++i) { avg += h[i]; if (h[i] > mx) mx = h[i]; }
  avg /= grid;
  const double bytes = static_cast<double>(p.nloads) * p.box_bytes * nprod;
  const double rows = static_cast<double>(p.nloads) * rows_per_box * nprod;
  printf("%-44s np %d st %2d grid %3d  box %3dB x %3d rows : %6.1f B/clk/SM  %5.2f clk/row  (avg %.0f clk, max %llu)\n", label, nprod, stages, grid,
         box_cols * 2, rows_per_box, bytes / avg, avg / rows, avg, mx);
  cudaFree(d);
  cudaFree(p.out);
  return 0;
}

int main() {
  for (int grid : {148, 16}) {
    run("load 128B x128 (16KB)", grid, 320, 64, 128, 0, CU_TENSOR_MAP_SWIZZLE_128B, 6, 1);
    run("load 128B x256 (32KB)", grid, 320, 64, 256, 0, CU_TENSOR_MAP_SWIZZLE_128B, 4, 1);
    run("load 128B x128 (16KB) 2 producers", grid, 320, 64, 128, 0, CU_TENSOR_MAP_SWIZZLE_128B, 4, 2);
    run("load 128B x128 (16KB) 3 producers", grid, 320, 64, 128, 0, CU_TENSOR_MAP_SWIZZLE_128B, 4, 3);
    run("load 128B x64  (8KB) 2 producers", grid, 320, 64, 64, 0, CU_TENSOR_MAP_SWIZZLE_128B, 6, 2);
    run("load 128B x64  (8KB) 4 producers", grid, 320, 64, 64, 0, CU_TENSOR_MAP_SWIZZLE_128B, 6, 4);
    run("load 128B x256 (32KB) 2 producers", grid, 320, 64, 256, 0, CU_TENSOR_MAP_SWIZZLE_128B, 3, 2);
  }
  return 0;
}
