// tma_probe.cu -- which of the TMA / mbarrier ingredients of me_frac3.cu this GPU + driver accepts, one variant per process
// (an illegal instruction poisons the context).  usage: tma_probe <variant>
//   0 mbarrier init + fence   1 + expect_tx(0) + try_wait   2 2-D tile load, map as __grid_constant__   3 4-D map
//   7 as 2 + fence.proxy.async after the mbarrier init   8 as 7 without .tile   9 as 8 at a 16-byte aligned x
//   10 / 11: 4-D box 32 x 9 without / with SWIZZLE_128B   12 / 13: box 80 x 5 without / with   (layout of the box in shared memory)
//   4 4-D map with SWIZZLE_128B   5 4-D map read from global memory   6 as 3, issued by every lane of a divergent branch
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

struct Maps { CUtensorMap m[2]; };

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void probe(const __grid_constant__ Maps maps, const CUtensorMap* gmaps, int variant, uint8_t* out)
{
  extern __shared__ unsigned char dyn[];
  unsigned char* base = (unsigned char*)(((uintptr_t)dyn + 1023) & ~(uintptr_t)1023);
  unsigned long long* bar = (unsigned long long*)(base + 8192);
  const int lane = threadIdx.x & 31;
  const uint32_t b = smem_u32(bar);
  if (lane == 0)
  {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (variant >= 7) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncwarp();
  if (variant == 0) { if (lane == 0) out[0] = 1; return; }
  const uint32_t bytes = variant == 1 ? 0u : (variant >= 12 ? 80u * 5u : (variant >= 10 ? 32u * 9u : 16u * 9u));
  if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
  __syncwarp();
  const uint32_t dst = smem_u32(base);
  if (variant == 8 && lane == 0)
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(&maps.m[0]), "r"(5), "r"(3), "r"(b) : "memory");
  if (variant == 9 && lane == 0)
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(&maps.m[0]), "r"(16), "r"(3), "r"(b) : "memory");
  if ((variant == 2 || variant == 7) && lane == 0)
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(&maps.m[0]), "r"(5), "r"(3), "r"(b) : "memory");
  if ((variant == 3 || variant == 4) && lane == 0)
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 ::"r"(dst), "l"(&maps.m[1]), "r"(5), "r"(3), "r"(2), "r"(1), "r"(b) : "memory");
  if (variant >= 10 && lane == 0)
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 ::"r"(dst), "l"(&maps.m[1]), "r"(16), "r"(3), "r"(2), "r"(1), "r"(b) : "memory");
  if (variant == 5 && lane == 0)
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 ::"r"(dst), "l"(gmaps + 1), "r"(5), "r"(3), "r"(2), "r"(1), "r"(b) : "memory");
  if (variant == 6 && (lane & 7) == 0)
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 ::"r"(dst + lane * 256), "l"(&maps.m[1]), "r"(5 + lane), "r"(3), "r"(2), "r"(1), "r"(b) : "memory");
  asm volatile(
    "{\n\t.reg .pred p;\n\t"
    "W:\n\t"
    "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
    "@p bra D;\n\t"
    "bra W;\n\t"
    "D:\n\t}" ::"r"(b), "r"(0) : "memory");
  for (int i = lane; i < 2048; i += 32) out[i] = base[i];
}

typedef CUresult (*Encode)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                           const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("variant %d: %s -> %s\n", variant, #x, cudaGetErrorString(e)); return 1; } } while (0)

int main(int argc, char** argv)
{
  const int variant = argc > 1 ? atoi(argv[1]) : 0;
  const int pitch = 640, ph = 400, planes = 16, slots = 2;
  const size_t plane = (size_t)pitch * ph, slot = plane * planes + 512;
  uint8_t* d;
  CK(cudaMalloc(&d, slot * slots));
  uint8_t* h = (uint8_t*)malloc(slot * slots);
  for (size_t i = 0; i < slot * slots; i++) h[i] = (uint8_t)(i * 2654435761u >> 13);
  CK(cudaMemcpy(d, h, slot * slots, cudaMemcpyHostToDevice));
  void* fn = NULL;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  Encode enc = (Encode)fn;
  Maps maps;
  {
    const cuuint64_t dim[2] = { (cuuint64_t)pitch, (cuuint64_t)ph }; const cuuint64_t str[1] = { (cuuint64_t)pitch };
    const cuuint32_t box[2] = { 16, 9 }, est[2] = { 1, 1 };
    CUresult r = enc(&maps.m[0], CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, dim, str, box, est, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r) { printf("encode 2d failed %d\n", (int)r); return 1; }
  }
  {
    const cuuint64_t dim[4] = { (cuuint64_t)pitch, (cuuint64_t)ph, (cuuint64_t)planes, (cuuint64_t)slots };
    const cuuint64_t str[3] = { (cuuint64_t)pitch, (cuuint64_t)plane, (cuuint64_t)slot };
    const cuuint32_t box[4] = { (cuuint32_t)(variant >= 12 ? 80 : (variant >= 10 ? 32 : 16)), (cuuint32_t)(variant >= 12 ? 5 : 9), 1, 1 }, est[4] = { 1, 1, 1, 1 };
    CUresult r = enc(&maps.m[1], CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, d, dim, str, box, est, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     (variant == 4 || variant == 11 || variant == 13) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r) { printf("encode 4d failed %d\n", (int)r); return 1; }
  }
  CUtensorMap* gm;
  CK(cudaMalloc(&gm, sizeof maps));
  CK(cudaMemcpy(gm, &maps, sizeof maps, cudaMemcpyHostToDevice));
  uint8_t* out;
  CK(cudaMalloc(&out, 8192));
  CK(cudaMemset(out, 0, 8192));
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
  probe<<<1, 32, 32768>>>(maps, gm, variant, out);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  uint8_t got[2048];
  CK(cudaMemcpy(got, out, 2048, cudaMemcpyDeviceToHost));
  if (variant >= 10)
  {
    // where did byte (r, c) of the box land?  hypotheses: dense rows of `iw` bytes; the same XOR-swizzled (chunk ^= (addr >> 7) & 7)
    const int iw = variant >= 12 ? 80 : 32, rows = variant >= 12 ? 5 : 9;
    int bad_dense = 0, bad_swz = 0;
    for (int r = 0; r < rows; r++)
      for (int c = 0; c < iw; c++)
      {
        const size_t src = slot * 1 + plane * 2 + (size_t)(3 + r) * pitch + 16 + c;
        const int off = r * iw + c, sw = off ^ ((off >> 3) & 0x70);
        if (got[off] != h[src]) bad_dense++;
        if (got[sw] != h[src]) bad_swz++;
      }
    printf("variant %d: box %d x %d: mismatches dense %d, dense + 128B swizzle %d (of %d)\n", variant, iw, rows, bad_dense, bad_swz, iw * rows);
    return 0;
  }
  if (variant < 2) { printf("variant %d: ok\n", variant); return 0; }
  // expected box: rows 3..11, bytes 5..20 of plane 2 of slot 1 (2-D: plane 0 of slot 0)
  int bad = 0, bad_swz = 0;
  for (int r = 0; r < 9; r++)
    for (int c = 0; c < 16; c++)
    {
      const bool two_d = variant == 2 || variant >= 7;
      const size_t src = (two_d ? 0 : slot * 1 + plane * 2) + (size_t)(3 + r) * pitch + (variant == 9 ? 16 : 5) + c;
      const int off = r * 16 + c;
      const int swz = off ^ ((off >> 3) & 0x70);
      if (got[off] != h[src]) bad++;
      if (swz < 144 && got[swz] != h[src]) bad_swz++;
    }
  printf("variant %d: ran; mismatches dense layout %d, 128B-swizzled layout %d (of 144)\n", variant, bad, bad_swz);
  return 0;
}
