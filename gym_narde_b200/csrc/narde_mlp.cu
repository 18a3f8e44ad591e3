// narde_mlp.cu -- afterstate scoring MLP (BASELINE config 5) on the 5th-gen tensor cores.
//
// Architecture = the reference's DecomposedDQN.forward(x) (train_deepq_pytorch.py:184-236) with
// state_size = 198:  x[K,198] -> Linear(198,256)+ReLU -> Linear(256,256)+ReLU -> Linear(256,576).
// One CTA owns a tile of 128 rows and runs all three layers back to back without leaving the SM:
//   * operands are bf16 in shared memory in the canonical K-major "interleave" (no-swizzle) UMMA
//     layout (8x8 core matrices of 128 B); activations are written there directly by the epilogue,
//     weights arrive pre-packed in that layout as contiguous 32 KB stages via cp.async.bulk (TMA
//     1-D bulk copy) completing on an mbarrier;
//   * tcgen05.mma (cta_group::1, kind::f16, M=128, N<=256, K=16) issued by one thread, fp32
//     accumulators in tensor memory (256 columns per 128x256 block);
//   * epilogue: tcgen05.ld (32 lanes x 32 columns per warp), + bias, ReLU, bf16, back to shared
//     memory as the next layer's A operand; the last layer streams fp32 Q-values to HBM.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/narde_b200.h"

namespace {

constexpr int kRows = 128;       // rows per CTA (UMMA M)
constexpr int kIn = 198;         // Box(198)
constexpr int kK = 256;          // padded K of every layer
constexpr int kH = 256;          // hidden width
constexpr int kOut = 576;        // move space (24*24)
constexpr int kKC = 64;          // K elements per weight stage
constexpr int kStages = 3;       // weight stages in flight
constexpr int kStageBytes = 256 * kKC * 2;  // 32 KB: a 256-row (N) x 64 (K) bf16 block
constexpr int kABytes = kRows * kK * 2;     // 64 KB activation tile
constexpr int kThreads = 512;      // 16 warps: TMEM lane quarter = warp & 3, column group = warp >> 2

// shared memory map (dynamic): [A0 | A1 | B stages | barriers]
constexpr int kSmemA0 = 0;
constexpr int kSmemA1 = kSmemA0 + kABytes;
constexpr int kSmemB = kSmemA1 + kABytes;
constexpr int kSmemBar = kSmemB + kStages * kStageBytes;
constexpr int kSmemBytes = kSmemBar + 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, SWIZZLE_NONE shared-memory matrix descriptor (cute/arch/mma_sm100_desc.hpp SmemDescriptor):
// start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48) | layout_type=0 [61,64)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
// instruction descriptor (InstrDescriptor): D=F32 [4,6)=1, A=BF16 [7,10)=1, B=BF16 [10,13)=1, K-major both,
// N>>3 at [17,23), M>>4 at [24,29)
__device__ __forceinline__ uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// byte offset of element (row r, k) in a K-major interleave tile of `rows` rows:
// core matrix = 8 rows x 16 B; row groups contiguous (SBO = 128 B), K chunks strided by LBO = rows*16 B
__device__ __forceinline__ uint32_t tile_off(int rows, int r, int k) {
  return (uint32_t)((k >> 3) * (rows * 16) + (r >> 3) * 128 + (r & 7) * 16 + (k & 7) * 2);
}

struct Layer {
  const uint8_t* w;   // packed stages: for each N block (<=256 rows), for each K chunk of 64: contiguous block
  const float* bias;
  int n_out;          // 256, 256, 576
};

// One GEMM block: D[128 x nb] = A[128 x 256] * Wblock^T, weights streamed through the stage ring.
// `it` is the running stage counter (ring position / mbarrier phase), shared by all layers.
__device__ __forceinline__ void gemm_block(uint32_t a_saddr, const uint8_t* wblk, int nb, uint32_t tmem_d, uint32_t bsm,
                                           uint32_t bar_full, uint32_t bar_empty, uint32_t bar_acc, uint32_t& it,
                                           uint32_t& acc_phase, int tid) {
  const int n_chunks = kK / kKC;              // 4
  const uint32_t stage_bytes = (uint32_t)nb * kKC * 2;
  if (tid == 0) {
    const uint32_t idesc = make_idesc(kRows, nb);
    // prologue: fill up to kStages stages
    int issued = 0;
    for (; issued < n_chunks && issued < kStages; issued++) {
      uint32_t s = (it + issued) % kStages, ph = ((it + issued) / kStages) & 1u;
      mbar_wait(bar_empty + 8 * s, ph ^ 1u);
      mbar_expect_tx(bar_full + 8 * s, stage_bytes);
      bulk_g2s(bsm + s * kStageBytes, wblk + (size_t)issued * stage_bytes, stage_bytes, bar_full + 8 * s);
    }
    for (int c = 0; c < n_chunks; c++) {
      uint32_t s = (it + c) % kStages, ph = ((it + c) / kStages) & 1u;
      mbar_wait(bar_full + 8 * s, ph);
      tc_fence_after();
#pragma unroll
      for (int kk = 0; kk < kKC / 16; kk++) {
        // A: k chunk index (c*64 + kk*16)/8 ; LBO = rows*16 (K direction), SBO = 128 (row groups)
        uint64_t ad = make_desc(a_saddr + (uint32_t)((c * kKC + kk * 16) >> 3) * (kRows * 16), kRows * 16, 128);
        uint64_t bd = make_desc(bsm + s * kStageBytes + (uint32_t)((kk * 16) >> 3) * (nb * 16), nb * 16, 128);
        umma(tmem_d, ad, bd, idesc, (c | kk) ? 1u : 0u);
      }
      umma_commit(bar_empty + 8 * s);          // frees the stage when these MMAs have read it
      if (issued < n_chunks) {                 // refill the ring
        uint32_t s2 = (it + issued) % kStages, ph2 = ((it + issued) / kStages) & 1u;
        mbar_wait(bar_empty + 8 * s2, ph2 ^ 1u);
        mbar_expect_tx(bar_full + 8 * s2, stage_bytes);
        bulk_g2s(bsm + s2 * kStageBytes, wblk + (size_t)issued * stage_bytes, stage_bytes, bar_full + 8 * s2);
        issued++;
      }
    }
    umma_commit(bar_acc);                      // accumulator complete
  }
  it += n_chunks;
  mbar_wait(bar_acc, acc_phase);
  acc_phase ^= 1u;
  tc_fence_after();
}

__global__ void __launch_bounds__(kThreads, 1)
k_mlp_forward(const float* __restrict__ x, int64_t rows, Layer L1, Layer L2, Layer L3, float* __restrict__ q) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t sbase = smem_u32(smem);
  const uint32_t a0 = sbase + kSmemA0, a1 = sbase + kSmemA1, bsm = sbase + kSmemB;
  const uint32_t bar_full = sbase + kSmemBar, bar_empty = bar_full + 8 * kStages, bar_acc = bar_empty + 8 * kStages;

  if (tid == 0) {
    for (int s = 0; s < kStages; s++) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_acc, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {  // one warp allocates 256 TMEM columns (128 lanes x 256 fp32)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&tmem_base_s)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;

  uint32_t it = 0, acc_phase = 0;
  for (int64_t tile = blockIdx.x; tile * kRows < rows; tile += gridDim.x) {
    const int64_t row0 = tile * kRows;
    // ---- stage X: fp32 [128,198] -> bf16 K-major interleave tile in A0 (K padded to 256 with zeros)
    for (int idx = tid; idx < kRows * (kK / 8); idx += kThreads) {
      int r = idx & (kRows - 1), kc = idx / kRows;  // consecutive threads -> consecutive rows (conflict-free stores)
      int64_t gr = row0 + r;
      __nv_bfloat16 h[8];
#pragma unroll
      for (int j = 0; j < 8; j++) {
        int k = kc * 8 + j;
        float v = (gr < rows && k < kIn) ? x[gr * kIn + k] : 0.0f;
        h[j] = __float2bfloat16(v);
      }
      *reinterpret_cast<uint4*>(smem + kSmemA0 + tile_off(kRows, r, kc * 8)) = *reinterpret_cast<uint4*>(h);
    }
    fence_async_smem();
    __syncthreads();

    // ---- layer 1: A0 -> A1, layer 2: A1 -> A0 ----
#pragma unroll 1
    for (int layer = 0; layer < 2; layer++) {
      const Layer& L = layer == 0 ? L1 : L2;
      const uint32_t a_in = layer == 0 ? a0 : a1;
      uint8_t* a_out = smem + (layer == 0 ? kSmemA1 : kSmemA0);
      gemm_block(a_in, L.w, kH, tmem, bsm, bar_full, bar_empty, bar_acc, it, acc_phase, tid);
      // epilogue: row = TMEM lane = (warp & 3) * 32 + lane; the 4 column groups (warp >> 2) take 64 columns each
      const int r = (warp & 3) * 32 + (tid & 31);
      const int cg = warp >> 2;
#pragma unroll 1
      for (int cb = cg * 2; cb < cg * 2 + 2; cb++) {
        uint32_t v[32];
        tmem_ld32(tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(cb * 32), v);
        const float4* b4 = reinterpret_cast<const float4*>(L.bias + cb * 32);
#pragma unroll
        for (int g = 0; g < 4; g++) {
          float4 ba = __ldg(b4 + 2 * g), bb = __ldg(b4 + 2 * g + 1);
          float f0 = __uint_as_float(v[g * 8 + 0]) + ba.x, f1 = __uint_as_float(v[g * 8 + 1]) + ba.y;
          float f2 = __uint_as_float(v[g * 8 + 2]) + ba.z, f3 = __uint_as_float(v[g * 8 + 3]) + ba.w;
          float f4 = __uint_as_float(v[g * 8 + 4]) + bb.x, f5 = __uint_as_float(v[g * 8 + 5]) + bb.y;
          float f6 = __uint_as_float(v[g * 8 + 6]) + bb.z, f7 = __uint_as_float(v[g * 8 + 7]) + bb.w;
          __nv_bfloat162 p0 = __floats2bfloat162_rn(fmaxf(f0, 0.0f), fmaxf(f1, 0.0f));
          __nv_bfloat162 p1 = __floats2bfloat162_rn(fmaxf(f2, 0.0f), fmaxf(f3, 0.0f));
          __nv_bfloat162 p2 = __floats2bfloat162_rn(fmaxf(f4, 0.0f), fmaxf(f5, 0.0f));
          __nv_bfloat162 p3 = __floats2bfloat162_rn(fmaxf(f6, 0.0f), fmaxf(f7, 0.0f));
          uint4 pk;
          pk.x = *reinterpret_cast<uint32_t*>(&p0);
          pk.y = *reinterpret_cast<uint32_t*>(&p1);
          pk.z = *reinterpret_cast<uint32_t*>(&p2);
          pk.w = *reinterpret_cast<uint32_t*>(&p3);
          *reinterpret_cast<uint4*>(a_out + tile_off(kRows, r, cb * 32 + g * 8)) = pk;
        }
      }
      tc_fence_before();
      fence_async_smem();
      __syncthreads();
      tc_fence_after();
    }

    // ---- layer 3: A0 -> Q[128,576] fp32 in HBM, N blocks of 256, 256, 64 ----
#pragma unroll 1
    for (int nb0 = 0; nb0 < kOut; nb0 += 256) {
      const int nb = kOut - nb0 < 256 ? kOut - nb0 : 256;
      gemm_block(a0, L3.w + (size_t)nb0 * kK * 2, nb, tmem, bsm, bar_full, bar_empty, bar_acc, it, acc_phase, tid);
      const int64_t gr = row0 + (warp & 3) * 32 + (tid & 31);
      const int cg = warp >> 2, per = nb >= 128 ? nb / 128 : 1;  // 32-column chunks per column group
#pragma unroll 1
      for (int cb = cg * per; cb < cg * per + per && cb * 32 < nb; cb++) {
        uint32_t v[32];
        tmem_ld32(tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(cb * 32), v);
        if (gr < rows) {
          float* dst = q + gr * kOut + nb0 + cb * 32;
          const float4* b4 = reinterpret_cast<const float4*>(L3.bias + nb0 + cb * 32);
#pragma unroll
          for (int j = 0; j < 8; j++) {
            float4 b = __ldg(b4 + j), o;
            o.x = __uint_as_float(v[4 * j]) + b.x;
            o.y = __uint_as_float(v[4 * j + 1]) + b.y;
            o.z = __uint_as_float(v[4 * j + 2]) + b.z;
            o.w = __uint_as_float(v[4 * j + 3]) + b.w;
            *reinterpret_cast<float4*>(dst + 4 * j) = o;
          }
        }
      }
      tc_fence_before();
      __syncthreads();
      tc_fence_after();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem));
}

bool g_mlp_attr_set = false;

}  // namespace

extern "C" {

// Packed weight layout (see gym_narde_b200/mlp.py:pack_weights): three layers back to back; per
// layer, per N block of <= 256 output rows, per K chunk of 64: nb x 64 bf16 in the K-major
// interleave layout (offset = (k/8)*(nb*16) + (n/8)*128 + (n%8)*16 + (k%8)*2), K padded to 256.
int narde_mlp_forward(const float* x, int64_t rows, const void* wpack, const float* bias, float* q, void* stream) {
  if (rows == 0) return 0;
  if (rows < 0 || !x || !wpack || !bias || !q) return -1;
  if ((((uintptr_t)wpack) & 15u) != 0 || (((uintptr_t)q) & 15u) != 0) return -1;
  if (!g_mlp_attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_mlp_forward, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (e != cudaSuccess) return (int)e;
    g_mlp_attr_set = true;
  }
  const uint8_t* w = (const uint8_t*)wpack;
  Layer L1 = {w, bias, kH};
  Layer L2 = {w + (size_t)kH * kK * 2, bias + kH, kH};
  Layer L3 = {w + (size_t)2 * kH * kK * 2, bias + 2 * kH, kOut};
  int64_t tiles = (rows + kRows - 1) / kRows;
  int grid = (int)(tiles < 148 ? tiles : 148);
  k_mlp_forward<<<grid, kThreads, kSmemBytes, (cudaStream_t)stream>>>(x, rows, L1, L2, L3, q);
  return (int)cudaGetLastError();
}

}  // extern "C"
