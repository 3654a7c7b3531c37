// Fused (flash) attention for the two full-sequence attentions on the hot path:
//   * CLIP ViT self-attention, S = 577, 16 heads x 64, no mask, scale 64^-0.5
//       (reference: CLIPVisionTower.forward -> transformers/models/clip/modeling_clip.py:261-345)
//   * LLaMA causal prefill attention over the paged KV cache, heads x 128, scale 128^-0.5, fp32 softmax
//       (reference: VCoder*LlavaLlamaForCausalLM.forward -> transformers/models/llama/modeling_llama.py:199-222; K11)
// One kernel: S = QK^T and O = PV on tensor cores (mma.sync m16n8k16, fp32 accumulate), online softmax in fp32
// registers with warp-shuffle row reductions, K/V tiles staged through XOR-swizzled shared memory with a
// double-buffered cp.async pipeline.  Scores never touch HBM: algorithmic traffic is Q,K,V in + O out.
// (Round-1 version on the legacy HMMA path; the tcgen05/TMEM version is the next step — see DESIGN.md.)
#include <cstdlib>
#include <string>

#include "common.h"
#include "kernels.h"
#include "ptx.cuh"

namespace vcb {
namespace {

constexpr int kBM = 64;  // query rows per CTA (4 warps x 16)
constexpr int kBN = 64;  // keys per tile (== KV page size on the paged path)
constexpr int kAttnThreads = 128;

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;  // src-size 0 => zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2,
                                                  uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void mma_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// smem tile: rows of DH halves, 16-byte chunks XOR-swizzled with (row & 7)
template <int DH>
__device__ __forceinline__ uint32_t tile_off(int row, int chunk) {
  return static_cast<uint32_t>(row * (DH * 2) + ((chunk ^ (row & 7)) << 4));
}

template <int DH, bool CAUSAL, bool PAGED>
__global__ void __launch_bounds__(kAttnThreads)
attn_prefill_kernel(AttnPrefillParams p) {
  constexpr int CH = DH / 8;                // 16B chunks per row
  constexpr int kTileBytes = kBN * DH * 2;  // one K or V tile
  extern __shared__ __align__(128) uint8_t smem[];
  // layout: Q tile | K[2] | V[2]
  const uint32_t sQ = smem_u32(smem);
  const uint32_t sK = sQ + kBM * DH * 2;
  const uint32_t sV = sK + 2 * kTileBytes;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int q0 = qt * kBM;
  const int S = p.seq_len;           // query rows per sequence
  const int P0 = PAGED ? p.q_pos0 : 0;  // position of query row 0: > 0 when a cached prefix is reused (f3)
  const int KV = P0 + S;             // keys visible to the last query row

  const __half* q_base = p.q + (static_cast<long>(b) * S) * p.q_stride + h * DH;
  // number of key tiles this query tile needs
  int n_tiles = (KV + kBN - 1) / kBN;
  if (CAUSAL) n_tiles = min(n_tiles, (P0 + q0 + kBM + kBN - 1) / kBN);

  auto load_kv = [&](int tile, int buf) {
    const __half* kb;
    const __half* vb;
    long stride;
    if (PAGED) {
      const int page = p.block_table[b * p.max_pages + tile];
      const long off = (static_cast<long>(page) * p.n_heads + h) * kBN * DH;
      kb = p.k_pool + off;
      vb = p.v_pool + off;
      stride = DH;
    } else {
      kb = p.k + (static_cast<long>(b) * S + tile * kBN) * p.kv_stride + h * DH;
      vb = p.v + (static_cast<long>(b) * S + tile * kBN) * p.kv_stride + h * DH;
      stride = p.kv_stride;
    }
#pragma unroll
    for (int i = 0; i < (kBN * CH) / kAttnThreads; ++i) {
      const int idx = tid + i * kAttnThreads;
      const int r = idx / CH, c = idx % CH;
      const bool ok = (tile * kBN + r) < KV;
      const long goff = ok ? (r * stride + c * 8) : 0;
      cp_async16(sK + buf * kTileBytes + tile_off<DH>(r, c), kb + goff, ok);
      cp_async16(sV + buf * kTileBytes + tile_off<DH>(r, c), vb + goff, ok);
    }
  };

  // ---- prologue: Q tile + first K/V tile
#pragma unroll
  for (int i = 0; i < (kBM * CH) / kAttnThreads; ++i) {
    const int idx = tid + i * kAttnThreads;
    const int r = idx / CH, c = idx % CH;
    const bool ok = (q0 + r) < S;
    cp_async16(sQ + tile_off<DH>(r, c), q_base + (ok ? (static_cast<long>(q0 + r) * p.q_stride + c * 8) : 0), ok);
  }
  load_kv(0, 0);
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();

  // Q fragments for this warp's 16 rows, all of DH
  uint32_t qf[DH / 16][4];
#pragma unroll
  for (int kk = 0; kk < DH / 16; ++kk) {
    const int row = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
    const int chunk = kk * 2 + (lane >> 4);
    ldmatrix_x4(sQ + tile_off<DH>(row, chunk), qf[kk][0], qf[kk][1], qf[kk][2], qf[kk][3]);
  }

  float o[DH / 8][4];
#pragma unroll
  for (int j = 0; j < DH / 8; ++j) { o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f; }
  float m_run[2] = {-INFINITY, -INFINITY};
  float l_run[2] = {0.f, 0.f};
  const float sc = p.scale * 1.4426950408889634f;  // scores are kept in the log2 domain
  const int qrow[2] = {q0 + warp * 16 + g, q0 + warp * 16 + g + 8};

  for (int tile = 0; tile < n_tiles; ++tile) {
    const int buf = tile & 1;
    if (tile + 1 < n_tiles) load_kv(tile + 1, buf ^ 1);
    cp_async_commit();

    // ---- S = Q K^T (16 x 64 per warp)
    float s[kBN / 8][4];
#pragma unroll
    for (int j = 0; j < kBN / 8; ++j) { s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f; }
    const uint32_t kt = sK + buf * kTileBytes;
#pragma unroll
    for (int kk = 0; kk < DH / 16; ++kk) {
#pragma unroll
      for (int jp = 0; jp < kBN / 16; ++jp) {
        uint32_t b0, b1, b2, b3;
        const int row = jp * 16 + (lane & 7) + (lane >> 4) * 8;
        const int chunk = kk * 2 + ((lane >> 3) & 1);
        ldmatrix_x4(kt + tile_off<DH>(row, chunk), b0, b1, b2, b3);
        mma_16816(s[2 * jp], qf[kk], b0, b1);
        mma_16816(s[2 * jp + 1], qf[kk], b2, b3);
      }
    }

    // ---- mask + online softmax (fp32)
    const int key0 = tile * kBN;
    const bool need_mask = (key0 + kBN > KV) || (CAUSAL && (key0 + kBN - 1 > P0 + q0 + warp * 16));
    float m_new[2] = {m_run[0], m_run[1]};
#pragma unroll
    for (int j = 0; j < kBN / 8; ++j) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float v = s[j][e] * sc;
        if (need_mask) {
          const int key = key0 + j * 8 + 2 * t + (e & 1);
          const int qr = qrow[e >> 1];
          if (key >= KV || (CAUSAL && key > P0 + qr)) v = -INFINITY;
        }
        s[j][e] = v;
        m_new[e >> 1] = fmaxf(m_new[e >> 1], v);
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      m_new[r] = fmaxf(m_new[r], __shfl_xor_sync(0xffffffffu, m_new[r], 1));
      m_new[r] = fmaxf(m_new[r], __shfl_xor_sync(0xffffffffu, m_new[r], 2));
    }
    float corr[2], m_use[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      m_use[r] = (m_new[r] == -INFINITY) ? 0.f : m_new[r];  // fully masked so far: keep exp2 finite
      corr[r] = exp2f(m_run[r] - m_use[r]);
      m_run[r] = m_new[r];
      l_run[r] *= corr[r];
    }
    float l_add[2] = {0.f, 0.f};
    uint32_t pf[kBN / 16][4];
#pragma unroll
    for (int j = 0; j < kBN / 8; ++j) {
      const float p0 = exp2f(s[j][0] - m_use[0]);
      const float p1 = exp2f(s[j][1] - m_use[0]);
      const float p2 = exp2f(s[j][2] - m_use[1]);
      const float p3 = exp2f(s[j][3] - m_use[1]);
      l_add[0] += p0 + p1;
      l_add[1] += p2 + p3;
      // C-fragment of two adjacent 8-key tiles == A-fragment of one 16-key block
      pf[j >> 1][(j & 1) * 2 + 0] = pack_half2(p0, p1);
      pf[j >> 1][(j & 1) * 2 + 1] = pack_half2(p2, p3);
    }
    l_run[0] += l_add[0];
    l_run[1] += l_add[1];
#pragma unroll
    for (int j = 0; j < DH / 8; ++j) {
      o[j][0] *= corr[0]; o[j][1] *= corr[0];
      o[j][2] *= corr[1]; o[j][3] *= corr[1];
    }

    // ---- O += P V
    const uint32_t vt = sV + buf * kTileBytes;
#pragma unroll
    for (int kk = 0; kk < kBN / 16; ++kk) {
#pragma unroll
      for (int jp = 0; jp < DH / 16; ++jp) {
        uint32_t b0, b1, b2, b3;
        const int row = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        const int chunk = jp * 2 + (lane >> 4);
        ldmatrix_x4_trans(vt + tile_off<DH>(row, chunk), b0, b1, b2, b3);
        mma_16816(o[2 * jp], pf[kk], b0, b1);
        mma_16816(o[2 * jp + 1], pf[kk], b2, b3);
      }
    }

    cp_async_wait<0>();
    __syncthreads();
  }

  // ---- finalise: divide by the row sums, store fp16
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
  }
  const float inv[2] = {l_run[0] > 0.f ? 1.f / l_run[0] : 0.f, l_run[1] > 0.f ? 1.f / l_run[1] : 0.f};
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    if (qrow[r] < S) {
      __half* orow = p.out + (static_cast<long>(b) * S + qrow[r]) * p.out_stride + h * DH;
#pragma unroll
      for (int j = 0; j < DH / 8; ++j) {
        *reinterpret_cast<uint32_t*>(orow + j * 8 + 2 * t) = pack_half2(o[j][2 * r] * inv[r], o[j][2 * r + 1] * inv[r]);
      }
    }
  }
}

template <int DH, bool CAUSAL, bool PAGED>
int launch_attn(const AttnPrefillParams& p, cudaStream_t stream) {
  constexpr int smem = (kBM * DH + 4 * kBN * DH) * 2;
  static bool attr = false;
  if (!attr) {
    VCB_CUDA(cudaFuncSetAttribute(attn_prefill_kernel<DH, CAUSAL, PAGED>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  smem));
    attr = true;
  }
  dim3 grid((p.seq_len + kBM - 1) / kBM, p.n_heads, p.batch);
  KernelScope scope(CAUSAL ? kCatAttnPrefill : kCatAttnVit, stream);
  attn_prefill_kernel<DH, CAUSAL, PAGED><<<grid, kAttnThreads, smem, stream>>>(p);
  VCB_CUDA(cudaGetLastError());
  return kOk;
}

}  // namespace

int attn_vit_legacy(const AttnPrefillParams& p, cudaStream_t stream) {
  VCB_CHECK(p.head_dim == 64, kErrInvalid, "attn_vit: head_dim must be 64");
  return launch_attn<64, false, false>(p, stream);
}

int attn_vit(const AttnPrefillParams& p, cudaStream_t stream) {
  static int use_tc5 = -1;  // 0 legacy mma.sync, 1 tcgen05 (one tile per CTA, two CTAs per SM), 2 tcgen05 pair schedule
  if (use_tc5 < 0) {
    const char* e = getenv("VCB_VIT_ATTN");
    const std::string v = e ? e : "";
    use_tc5 = v == "legacy" ? 0 : (v == "tc5" ? 1 : 2);  // default: pair schedule (fastest measured)
  }
  const bool packed = p.q_stride == p.kv_stride && p.k == p.q + p.n_heads * 64 && p.v == p.q + 2 * p.n_heads * 64;
  if (use_tc5 == 2 && packed && p.q_stride % 8 == 0) return attn_vit_pair(p, stream);
  if (use_tc5 && packed && p.q_stride % 8 == 0) return attn_vit_tc5(p, stream);
  return attn_vit_legacy(p, stream);
}

int attn_llama_prefill(const AttnPrefillParams& p, cudaStream_t stream) {
  VCB_CHECK(p.head_dim == 128, kErrInvalid, "attn_llama_prefill: head_dim must be 128");
  VCB_CHECK(p.page_tokens == kBN, kErrInvalid, "attn_llama_prefill: KV page must hold 64 tokens");
  // VCB_PREFILL_ATTN=legacy: the mma.sync / cp.async kernel of this file (A/B, and the fallback for unaligned buffers);
  // read on every call so that one test process can compare the two
  const char* e = getenv("VCB_PREFILL_ATTN");
  const bool legacy = e != nullptr && std::string(e) == "legacy";
  const bool aligned = p.q_stride % 8 == 0 && (reinterpret_cast<uintptr_t>(p.q) & 15) == 0 && p.out_stride % 8 == 0 &&
                       (reinterpret_cast<uintptr_t>(p.out) & 15) == 0;
  if (!legacy && aligned) return attn_llama_prefill_tc5(p, stream);
  return launch_attn<128, true, true>(p, stream);
}

}  // namespace vcb
