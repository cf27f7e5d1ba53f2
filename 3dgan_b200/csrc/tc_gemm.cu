// Hand-written sm_100a implicit-GEMM kernels: tcgen05.mma (bf16 x bf16 -> fp32 in TMEM) fed by TMA
// through a multi-stage mbarrier pipeline, warp-specialised (1 TMA warp, 1 MMA warp, 16 epilogue
// warps: four per TMEM lane quarter).  Replaces the cuDNN/cuBLAS calls TensorFlow made for the reference's
// tf.nn.conv2d / conv2d_transpose / matmul call sites (ops/layers.py:57,101,142;
// hem/ops/layers.py:61,118,189) and their autodiff gradients.
#include "tc_gemm.cuh"
#include "ptx.cuh"
#include "epilogue.cuh"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

namespace b200 {

namespace {

constexpr int kMaxThreads = 64 + 32 * 16;   // warp0 TMA, warp1 MMA, then 4..16 epilogue warps (launch parameter)
constexpr int kABytes = kTileM * kBlockK * 2;   // 16 KiB: 128 rows x 128 B

// one contiguous piece of a CTA pair's range that lies inside a single (phase, N tile, pixel group) item
struct SkSeg {
  uint16_t ntile, it0, len, first_pair;   // K-iterations [it0, it0 + len) of the item's (tap, K chunk) sequence;
                                          // finishing && it0 > 0: partials of pairs [first_pair, own) must be added
  uint32_t group;                         // pixel-tile group of the item
  uint8_t phase, finishing, pad_[2];      // finishing: contains the item's last iteration -> runs the item's epilogue
};
constexpr int kMaxSkSegs = 44;

struct PipeSmem {
  uint64_t full[8];
  uint64_t empty[8];
  uint64_t tmem_full;
  uint64_t tmem_empty;
  uint64_t acc_full[2];        // persistent kernel: one pair of barriers per TMEM accumulator buffer
  uint64_t acc_empty[2];
  uint32_t tmem_base;
  int nseg;                    // persistent kernel: pieces of this pair, in processing order
  long long trace[6];          // B200GAN_GEMM_TRACE: clock stamps of CTA (0,0,0) (debug)
  alignas(16) float bias[256]; // this N tile's bias row (epilogue)
  SkSeg segs[kMaxSkSegs];
};
static_assert(sizeof(PipeSmem) <= 2048, "PipeSmem must fit the slack pick_stages leaves");

__device__ __forceinline__ uint8_t* align1024(uint8_t* p) {
  uintptr_t a = reinterpret_cast<uintptr_t>(p);
  return reinterpret_cast<uint8_t*>((a + 1023) & ~uintptr_t(1023));
}

}  // namespace

// =============================================================================================
// Tap GEMM (conv fprop / dgrad phases / transposed-conv forward / dense)
// =============================================================================================
// kCluster == 2: the two CTAs of a cluster own adjacent pixel tiles (same N tile, same phase); each loads
// its own A tile and HALF of the shared B tile, multicast into both CTAs' shared memory, so the L2->SM
// traffic per K chunk drops from A+B to A+B/2.  Stage release (tcgen05.commit) is multicast to both CTAs'
// empty barriers because either producer writes into both CTAs.
// p.cluster_y == 2 (with two pixel tiles per CTA): the cluster is 2 x 2 -- the CTAs of the two N tiles that
// share the same pixel tiles each load ONE of the two A tiles and multicast it to the other, so a CTA
// pulls A/2 + B/2 from L2 for an (256 x bn_tile) output block.  These GEMMs are bound by the L2->SM
// fill rate (about 12 TB/s chip-wide), so halving the operand traffic is what raises tensor-pipe use.
template <int kCluster, bool kSimple>
__global__ void __launch_bounds__(kMaxThreads, 1) tapgemm_kernel(const __grid_constant__ TapGemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  griddep_launch_dependents();
  const uint32_t crank = kCluster > 1 ? cluster_ctarank() : 0;
  // cluster dims (2, cluster_y, 1): rank = x + 2*y.  cx pairs pixel tiles (shares B), cyi pairs N tiles (shares A)
  const uint32_t cx = crank & 1, cyi = crank >> 1;
  const bool share_a = kCluster > 1 && p.cluster_y == 2;
  const uint16_t b_mask = (uint16_t)(0x3u << (2 * cyi));
  const uint16_t a_mask = (uint16_t)((1u << cx) | (1u << (cx + 2)));

  const int b_bytes = p.bn_tile * kBlockK * 2;
  const int a_bytes = p.dual * kABytes;         // p.dual (1|2) pixel tiles per CTA share one B tile
  // p.merge_tail: the narrow tail chunk of a tap rides in the stage of the tap's last full chunk
  // (its own small area behind the full tiles), so it does not cost a pipeline slot of its own
  const int tail_area = p.merge_tail ? (p.dual * kTileM + p.bn_tile) * 32 : 0;
  const int stage_bytes = a_bytes + b_bytes + tail_area;
  PipeSmem* ps = reinterpret_cast<PipeSmem*>(smem + (size_t)p.stages * stage_bytes);

  // CTA -> p.dual consecutive pixel tiles (first pixel of each), phase, N offset.  p.splits > 1 (layers with only a
  // few output tiles but long reductions: the 1x1 .. 8x8 maps of pix2pix's U-Net, hem/models/pix2pix.py:187-227):
  // blockIdx.z also carries a K split -- each CTA runs a slice of the phase's (tap, K chunk) sequence and adds its
  // partial tile into an fp32 workspace with red.global.add; a finalize kernel applies the epilogue.
  const int nsplit = p.splits > 1 ? p.splits : 1;
  const int phase = blockIdx.z % p.nphases;
  const int split = blockIdx.z / p.nphases;
  int pw0[2], ph0[2], pn0[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    int t = blockIdx.x * p.dual + (i < p.dual ? i : 0);
    const int tw = t % p.tiles_w; t /= p.tiles_w;
    const int th = t % p.tiles_h; t /= p.tiles_h;
    pw0[i] = tw * p.bw; ph0[i] = th * p.bh; pn0[i] = t * p.bn;
  }
  const int n0 = blockIdx.y * p.bn_tile;
  const int ext_w = p.phase_ext_w[phase], ext_h = p.phase_ext_h[phase];
  const int tail_row_bytes = p.tail_mode == 1 ? 32 : 64;      // bytes per row of a narrow tail box
  // tiles outside this phase's extent / past the last tile load zeros (TMA OOB) and skip their stores

  const int tap_begin = p.phase_tap_begin[phase];
  const int ntaps = p.phase_tap_begin[phase + 1] - tap_begin;
  const int kloops = p.merge_tail ? p.kchunks - 1 : p.kchunks;     // pipeline iterations per tap
  const int iters_all = ntaps * kloops;
  const int it_begin = (int)((long long)iters_all * split / nsplit);
  const int it_end = (int)((long long)iters_all * (split + 1) / nsplit);
  const int iters = it_end - it_begin;

  const bool tr = p.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
  if (tr && threadIdx.x == 0) ps->trace[0] = clock64();
  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&p.tmA);
    tma_prefetch_desc(&p.tmB);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(smem_u32(&ps->full[s]), 1);
      mbar_init(smem_u32(&ps->empty[s]), kCluster + (share_a ? 1 : 0));   // every CTA a producer here writes into
    }
    mbar_init(smem_u32(&ps->tmem_full), 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    if (p.dual == 2) tmem_alloc<2 * kTmemCols>(smem_u32(&ps->tmem_base));
    else tmem_alloc<kTmemCols>(smem_u32(&ps->tmem_base));
  }
  tc_fence_before();
  __syncthreads();
  if (kCluster > 1) cluster_sync_all();       // peer barriers are initialised before anything lands on them
  tc_fence_after();
  const uint32_t tmem = ps->tmem_base;
  griddep_wait();      // everything above touched only shared memory / TMEM / kernel parameters

  if (warp == 0) {
    if (elect_one()) {
      // ------------------------------------------------------------------ TMA producer
      int base[2][5];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        base[i][0] = 0;
#pragma unroll
        for (int d = 0; d < 4; ++d)
          base[i][d + 1] = pw0[i] * p.a_mul[0][d] + ph0[i] * p.a_mul[1][d] + pn0[i] * p.a_mul[2][d];
      }
      int s = 0;
      uint32_t par = 0;
      const uint32_t full0 = smem_u32(&ps->full[0]), empty0 = smem_u32(&ps->empty[0]);
      const uint32_t smem0 = smem_u32(smem);
      // L2 prefetch iterator, p.l2_prefetch K blocks ahead of the loads
      int ptp = 0, pkc = 0;
      auto prefetch_next = [&]() {
        if (ptp < ntaps) {
          const int tap = tap_begin + ptp;
          int c[5];
          c[0] = pkc * kBlockK;
#pragma unroll
          for (int d = 0; d < 4; ++d) c[d + 1] = base[0][d + 1] + p.tap_a_off[tap][d];
          tma_prefetch_nd(p.a_rank, &p.tmA, c);
          if (p.dual == 2) {
#pragma unroll
            for (int d = 0; d < 4; ++d) c[d + 1] = base[1][d + 1] + p.tap_a_off[tap][d];
            tma_prefetch_nd(p.a_rank, &p.tmA, c);
          }
          int cb[2] = {pkc * kBlockK, p.tap_b_row[tap] + n0 + (kCluster > 1 ? (int)cx * (p.bn_tile / 2) : 0)};
          tma_prefetch_nd(2, &p.tmB, cb);
          if (++pkc == p.kchunks) { pkc = 0; ++ptp; }
        }
      };
      const bool l2pf = p.l2_prefetch && nsplit == 1;
      if (l2pf) for (int i = 0; i < p.l2_prefetch; ++i) prefetch_next();
      int tp = it_begin / kloops, kc = it_begin - tp * kloops;
      int c0[5], c1[5], brow = 0;
      bool new_tap = true;
      {
        for (int it = it_begin; it < it_end; ++it) {
          if (new_tap) {
            const int tap = tap_begin + tp;
#pragma unroll
            for (int d = 0; d < 4; ++d) {
              c0[d + 1] = base[0][d + 1] + p.tap_a_off[tap][d];
              c1[d + 1] = base[1][d + 1] + p.tap_a_off[tap][d];
            }
            brow = p.tap_b_row[tap] + n0;
            new_tap = false;
          }
          if (l2pf) prefetch_next();
          mbar_wait(empty0 + 8 * s, par ^ 1);
          const uint32_t full = full0 + 8 * s;
          const uint32_t a_dst = smem0 + s * stage_bytes;
          c0[0] = kc * kBlockK;
          c1[0] = kc * kBlockK;
          const bool tail = !p.merge_tail && p.tail_mode != 0 && kc == p.kchunks - 1;
          const bool with_tail = p.merge_tail && kc == kloops - 1;
          // the tail chunk of a tap moves only its 16 / 32 valid channels (narrow box, 32B / 64B swizzle)
          const int row_bytes = tail ? tail_row_bytes : kBlockK * 2;
          const int a_tile = kTileM * row_bytes;
          const void* mA = tail ? (const void*)&p.tmA_tail : (const void*)&p.tmA;
          const void* mB = tail ? (const void*)&p.tmB_tail : (const void*)&p.tmB;
          mbar_arrive_expect_tx(full, p.dual * a_tile + p.bn_tile * row_bytes + (with_tail ? tail_area : 0));
          if (share_a) {
            tma_load_nd_mc(p.a_rank, a_dst + cyi * a_tile, mA, full, cyi ? c1 : c0, a_mask);
          } else {
            tma_load_nd(p.a_rank, a_dst, mA, full, c0);
            if (p.dual == 2) tma_load_nd(p.a_rank, a_dst + a_tile, mA, full, c1);
          }
          if (kCluster == 1) {
            tma_load_2d(a_dst + p.dual * a_tile, mB, full, kc * kBlockK, brow);
          } else {
            const int half_rows = p.bn_tile / 2;
            tma_load_2d_mc(a_dst + p.dual * a_tile + cx * half_rows * row_bytes, mB, full, kc * kBlockK,
                           brow + cx * half_rows, b_mask);
          }
          if (with_tail) {
            // 16-wide tail boxes (rows of 32 B) behind the full tiles of this stage
            const uint32_t t_dst = a_dst + a_bytes + b_bytes;
            const int t_tile = kTileM * 32;
            c0[0] = (kc + 1) * kBlockK;
            c1[0] = (kc + 1) * kBlockK;
            if (share_a) {
              tma_load_nd_mc(p.a_rank, t_dst + cyi * t_tile, &p.tmA_tail, full, cyi ? c1 : c0, a_mask);
            } else {
              tma_load_nd(p.a_rank, t_dst, &p.tmA_tail, full, c0);
              if (p.dual == 2) tma_load_nd(p.a_rank, t_dst + t_tile, &p.tmA_tail, full, c1);
            }
            if (kCluster == 1) {
              tma_load_2d(t_dst + p.dual * t_tile, &p.tmB_tail, full, (kc + 1) * kBlockK, brow);
            } else {
              const int half_rows = p.bn_tile / 2;
              tma_load_2d_mc(t_dst + p.dual * t_tile + cx * half_rows * 32, &p.tmB_tail, full, (kc + 1) * kBlockK,
                             brow + cx * half_rows, b_mask);
            }
          }
          if (++s == p.stages) { s = 0; par ^= 1; }
          if (++kc == kloops) { kc = 0; ++tp; new_tap = true; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (elect_one()) {
      // ------------------------------------------------------------------ MMA issuer
      const uint32_t idesc = make_idesc_bf16(kTileM, p.bn_tile, 0, 0);
      // the last K chunk of a tap is zero-filled beyond K: issue only the 16-wide steps that hold data
      const int tail_steps = (p.k_total - (p.kchunks - 1) * kBlockK + 15) / 16;
      // this loop is the critical path of the kernel (one thread feeds the tensor pipe): no divisions,
      // descriptors advanced by adds, barrier addresses by a stage counter
      const uint32_t smem0 = smem_u32(smem);
      const uint64_t adesc0 = make_smem_desc_sw128(smem0, 16, 1024);
      const uint64_t bdesc0 = make_smem_desc_sw128(smem0 + a_bytes, 16, 1024);
      const uint32_t desc_step = (uint32_t)stage_bytes >> 4;
      const uint32_t full0 = smem_u32(&ps->full[0]), empty0 = smem_u32(&ps->empty[0]);
      const bool dual = p.dual == 2;
      int s = 0, kc = it_begin % kloops;
      uint32_t par = 0, acc = 0;
      if (tr) ps->trace[1] = clock64();
      for (int it = 0; it < iters; ++it) {
        mbar_wait(full0 + 8 * s, par);
        if (tr && it == 0) ps->trace[2] = clock64();
        tc_fence_after();
        const uint64_t adesc = adesc0 + (uint64_t)(desc_step * s);
        const uint64_t bdesc = bdesc0 + (uint64_t)(desc_step * s);
        if (p.merge_tail) {
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {
            umma_bf16(tmem, adesc + 2 * k, bdesc + 2 * k, idesc, acc);
            if (dual) umma_bf16(tmem + kTmemCols, adesc + (kABytes >> 4) + 2 * k, bdesc + 2 * k, idesc, acc);
            acc = 1;
          }
          if (kc == kloops - 1) {
            // the tap's 16-wide tail (SWIZZLE_32B tiles behind the full tiles of this stage): one K step
            const uint32_t tb0 = smem0 + s * stage_bytes + a_bytes + b_bytes;
            const uint32_t t_tile = kTileM * 32;
            umma_bf16(tmem, make_smem_desc(tb0, 16, 256, 6), make_smem_desc(tb0 + p.dual * t_tile, 16, 256, 6), idesc, 1);
            if (dual)
              umma_bf16(tmem + kTmemCols, make_smem_desc(tb0 + t_tile, 16, 256, 6),
                        make_smem_desc(tb0 + p.dual * t_tile, 16, 256, 6), idesc, 1);
          }
        } else if (kc != p.kchunks - 1 || p.tail_mode == 0) {
          const int nsteps = (kc != p.kchunks - 1) ? kBlockK / 16 : tail_steps;
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {
            if (k < nsteps) {
              // +32 bytes per 16-element K step inside the 128-byte swizzle row (addr field is >>4);
              // the second pixel tile (A + 16 KiB) accumulates into TMEM columns [256, 256 + N)
              umma_bf16(tmem, adesc + 2 * k, bdesc + 2 * k, idesc, acc);
              if (dual) umma_bf16(tmem + kTmemCols, adesc + (kABytes >> 4) + 2 * k, bdesc + 2 * k, idesc, acc);
              acc = 1;
            }
          }
        } else {
          // narrow tail tile: rows of 32 B (SWIZZLE_32B, 8-row groups 256 B apart) or 64 B (SWIZZLE_64B, 512 B)
          const uint32_t base = smem0 + s * stage_bytes;
          const uint32_t a_tile = kTileM * tail_row_bytes;
          const uint32_t layout = p.tail_mode == 1 ? 6u : 4u;
          const uint32_t sbo = 8 * tail_row_bytes;
          const uint64_t ta0 = make_smem_desc(base, 16, sbo, layout);
          const uint64_t ta1 = make_smem_desc(base + a_tile, 16, sbo, layout);
          const uint64_t tb = make_smem_desc(base + p.dual * a_tile, 16, sbo, layout);
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            if (k < tail_steps) {
              umma_bf16(tmem, ta0 + 2 * k, tb + 2 * k, idesc, acc);
              if (dual) umma_bf16(tmem + kTmemCols, ta1 + 2 * k, tb + 2 * k, idesc, acc);
              acc = 1;
            }
          }
        }
        if (kCluster == 1) umma_commit(empty0 + 8 * s);
        else umma_commit_mc(empty0 + 8 * s, share_a ? (uint16_t)(b_mask | a_mask) : b_mask);
        if (++kc == kloops) kc = 0;
        if (++s == p.stages) { s = 0; par ^= 1; }
      }
      umma_commit(smem_u32(&ps->tmem_full));
      if (tr) ps->trace[3] = clock64();
    }
    __syncwarp();
  } else {
    // -------------------------------------------------------------------- epilogue (4 warps)
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int cg = (warp - 2) >> 2;         // column group: warps sharing a quarter split the column chunks
    const int ncg = ((int)(blockDim.x >> 5) - 2) >> 2;
    const int r = q * 32 + lane;            // tile row == TMEM lane
    const int iw = r % p.bw;
    const int ih = (r / p.bw) % p.bh;
    const int in = r / (p.bw * p.bh);
    EpilogueArgs ea;
    ea.bias = p.bias; ea.act = p.act; ea.leak = p.leak; ea.mask_src = p.mask_src;
    ea.mask_kind = p.mask_kind; ea.alpha = p.alpha; ea.out = p.out; ea.out_f32 = p.out_f32;
    ea.accumulate = p.accumulate; ea.ncols = p.ncols; ea.pipelined = p.epi_pipe;
    ea.mask_bits = p.mask_bits; ea.bits_out = p.bits_out; ea.bits_pitch = p.bits_pitch; ea.row_elems = p.row_elems;
    ea.stage_row = 0; ea.stage_bits = 0; ea.stage_col0 = 0;
    ea.bias_smem = 0; ea.bias_col0 = n0;
    ea.partial_row = nullptr; ea.partial_n = 0; ea.partial_stride = 0;
    if (p.bias) {
      for (int i = (int)threadIdx.x - 64; i < p.bn_tile; i += (int)blockDim.x - 64)
        ps->bias[i] = (n0 + i < p.ncols) ? __ldg(p.bias + n0 + i) : 0.f;
      named_barrier(2, (int)blockDim.x - 64);
      ea.bias_smem = smem_u32(&ps->bias[0]);
    }

    const int total_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
    bool row_ok[2];
    long long off[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int pw = pw0[i] + iw, ph = ph0[i] + ih, pn = pn0[i] + in;
      const bool tile_ok = i < p.dual && (int)(blockIdx.x * p.dual + i) < total_tiles;
      row_ok[i] = tile_ok && pw < ext_w && ph < ext_h && pn < p.ext_n;
      off[i] = p.phase_o_off[phase] + (long long)pn * p.o_sn + (long long)ph * p.o_sh + (long long)pw * p.o_sw;
      if (row_ok[i] && cg == 0) epilogue_prefetch_mask(ea, off[i], n0, p.bn_tile);
    }
    if (p.b_prefetch && blockIdx.x < kCluster) {
      // weights are read by every pixel-tile CTA in lockstep; when they are not L2-resident each K chunk is
      // a chip-wide HBM miss on the critical path.  The first cluster's idle epilogue threads pull this N
      // tile's weight rows into L2 while the pipeline starts.
      const int rows = p.bn_tile / kCluster;
      const int row0 = n0 + (int)cx * rows;
      const int lines = (p.k_total * 2 + 127) >> 7;
      const int per_tap = rows * lines;
      const int total = ntaps * per_tap;
      const char* base = reinterpret_cast<const char*>(p.b_base);
      for (int i = (int)threadIdx.x - 64; i < total; i += (int)blockDim.x - 64) {
        const int tp = i / per_tap, rem = i - tp * per_tap;
        const int rr = rem / lines, ln = rem - rr * lines;
        const int row = p.tap_b_row[tap_begin + tp] + row0 + rr;
        if (row < p.b_rows_total)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(base + (size_t)row * p.b_pitch_bytes + ln * 128));
      }
    }
    mbar_wait(smem_u32(&ps->tmem_full), 0);
    if (tr && threadIdx.x == 64) ps->trace[4] = clock64();
    tc_fence_after();
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      if (i >= p.dual) break;
      const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16) + i * kTmemCols;
      if (p.tma_store) {
        // all MMAs have completed (tmem_full), so the pipeline stages are idle: tile i is staged at
        // smem + i * 128 rows and leaves through one bulk tensor store
        ea.stage_row = smem_u32(smem) + (uint32_t)((i * kTileM + r) * p.stage_pitch);
        ea.stage_col0 = n0;
      }
      epilogue_row<kSimple>(ea, trow, off[i], row_ok[i], n0, cg * 16, ncg * 16, min(p.bn_tile, p.ncols - n0));
    }
    if (p.tma_store) {
      fence_proxy_async_smem();
      named_barrier(1, (int)blockDim.x - 64);
      if (threadIdx.x == 64) {
        for (int i = 0; i < p.dual; ++i)
          if ((int)(blockIdx.x * p.dual + i) < total_tiles)
            tma_store_4d(&p.tmOut[phase], smem_u32(smem) + (uint32_t)(i * kTileM * p.stage_pitch), n0, pw0[i], ph0[i],
                         pn0[i]);
        bulk_commit();
        bulk_wait_read0();       // shared memory must outlive the copy engine's reads
      }
    }
  }

  if (tr && threadIdx.x == 64) ps->trace[5] = clock64();
  tc_fence_before();
  __syncthreads();
  if (tr && threadIdx.x == 0) {
    const long long t0 = ps->trace[0];
    printf("tapgemm grid(%d,%d,%d) iters %d: setup_done %lld first_data %lld mma_issued %lld epi_start %lld epi_end(warp2) %lld all_done %lld\n",
           gridDim.x, gridDim.y, gridDim.z, iters, ps->trace[1] - t0, ps->trace[2] - t0, ps->trace[3] - t0, ps->trace[4] - t0,
           ps->trace[5] - t0, (long long)clock64() - t0);
  }
  if (kCluster > 1) cluster_sync_all();       // the peer may still be signalling our barriers
  if (warp == 1) {
    if (p.dual == 2) tmem_dealloc<2 * kTmemCols>(tmem);
    else tmem_dealloc<kTmemCols>(tmem);
  }
}

// =============================================================================================
// Tap GEMM, 2-CTA form (tcgen05 cta_group::2)
// =============================================================================================
// The cluster pair runs UMMA instructions with M = 256: rows 0..127 are the leader's pixel tile, rows 128..255
// the peer's, the accumulators land in each CTA's own TMEM; each CTA keeps only HALF of the B tile in shared
// memory (rows [rank * N/2, (rank+1) * N/2)), the tensor core reads both halves.  Per 64-wide K block a CTA
// therefore receives 2 A tiles + B/2 = 45 KB instead of 58 KB: the 1-CTA kernel is bound by the shared-memory
// fill rate (about 64 B/clk/SM; 58 KB per 832 MMA cycles), this one is not.  Protocol (as CUTLASS's
// PipelineTmaUmmaAsync): both producers wait on their own empty barrier and issue cta_group::2 TMA loads that
// complete bytes on the LEADER's full barrier; the leader's producer alone arms it with the pair's byte count;
// the leader's MMA thread issues every MMA and releases stages / publishes the accumulators with multicast commits.
// Requires two pixel tiles per CTA.
template <bool kSimple>
__global__ void __launch_bounds__(kMaxThreads, 1) tapgemm2sm_kernel(const __grid_constant__ TapGemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // cluster dims (2,1,1): rank 0 = leader.  p.quad: cluster (2,2,1) = two pairs on adjacent N tiles (ranks 0,1 and 2,3,
  // even rank = leader); they need the same pixel tiles, so each CTA loads ONE of its two A tiles and multicasts it to
  // its twin in the other pair: 29 KB instead of 45 KB of L2 reads per CTA and pipeline iteration
  const uint32_t qrank = cluster_ctarank();
  const uint32_t crank = qrank & 1;                  // rank within the CTA pair
  const bool leader = crank == 0;
  const bool quad = p.quad != 0;
  const uint16_t pair_mask = (uint16_t)(0x3u << (qrank & 2));
  const uint16_t twin_mask = (uint16_t)((1u << qrank) | (1u << (qrank ^ 2)));

  const int half_rows = p.bn_tile / 2;
  const int b_bytes = half_rows * kBlockK * 2;        // this CTA's half of the B tile
  const int a_bytes = 2 * kABytes;                    // two pixel tiles per CTA
  const int tail_area = p.merge_tail ? (2 * kTileM + half_rows) * 32 : 0;
  const int stage_bytes = a_bytes + b_bytes + tail_area;
  PipeSmem* ps = reinterpret_cast<PipeSmem*>(smem + (size_t)p.stages * stage_bytes);

  const int phase = blockIdx.z;
  int pw0[2], ph0[2], pn0[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    int t = blockIdx.x * 2 + i;
    const int tw = t % p.tiles_w; t /= p.tiles_w;
    const int th = t % p.tiles_h; t /= p.tiles_h;
    pw0[i] = tw * p.bw; ph0[i] = th * p.bh; pn0[i] = t * p.bn;
  }
  const int n0 = blockIdx.y * p.bn_tile;
  const int ext_w = p.phase_ext_w[phase], ext_h = p.phase_ext_h[phase];
  const int tap_begin = p.phase_tap_begin[phase];
  const int ntaps = p.phase_tap_begin[phase + 1] - tap_begin;
  const int kloops = p.merge_tail ? p.kchunks - 1 : p.kchunks;
  const int iters = ntaps * kloops;
  const int tail_row_bytes = p.tail_mode == 1 ? 32 : 64;      // bytes per row of a narrow (not merged) tail box

  const bool tr = p.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
  if (tr && threadIdx.x == 0) ps->trace[0] = clock64();
  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&p.tmA);
    tma_prefetch_desc(&p.tmB);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(smem_u32(&ps->full[s]), 1);
      mbar_init(smem_u32(&ps->empty[s]), quad ? 2 : 1);    // quad: both pairs' MMAs must be done with the stage
    }
    mbar_init(smem_u32(&ps->tmem_full), 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc_2sm<2 * kTmemCols>(smem_u32(&ps->tmem_base));
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();       // both CTAs' barriers and TMEM exist before any load / MMA touches them
  tc_fence_after();
  const uint32_t tmem = ps->tmem_base;

  if (warp == 0) {
    if (elect_one()) {
      // ------------------------------------------------------------------ TMA producer (both CTAs)
      int base[2][5];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        base[i][0] = 0;
#pragma unroll
        for (int d = 0; d < 4; ++d)
          base[i][d + 1] = pw0[i] * p.a_mul[0][d] + ph0[i] * p.a_mul[1][d] + pn0[i] * p.a_mul[2][d];
      }
      int s = 0;
      uint32_t par = 0;
      const uint32_t full0 = smem_u32(&ps->full[0]), empty0 = smem_u32(&ps->empty[0]);
      const uint32_t smem0 = smem_u32(smem);
      const int a_tile = kABytes, t_tile = kTileM * 32;
      for (int tp = 0; tp < ntaps; ++tp) {
        const int tap = tap_begin + tp;
        int c0[5], c1[5];
#pragma unroll
        for (int d = 0; d < 4; ++d) {
          c0[d + 1] = base[0][d + 1] + p.tap_a_off[tap][d];
          c1[d + 1] = base[1][d + 1] + p.tap_a_off[tap][d];
        }
        const int brow = p.tap_b_row[tap] + n0 + (int)crank * half_rows;
        for (int kc = 0; kc < kloops; ++kc) {
          mbar_wait(empty0 + 8 * s, par ^ 1);
          const uint32_t full = full0 + 8 * s;
          const uint32_t a_dst = smem0 + s * stage_bytes;
          const bool with_tail = p.merge_tail && kc == kloops - 1;
          // a tail chunk that is not merged moves only its 16 / 32 valid channels (narrow box, 32B / 64B swizzle)
          const bool tail = !p.merge_tail && p.tail_mode != 0 && kc == p.kchunks - 1;
          c0[0] = kc * kBlockK;
          c1[0] = kc * kBlockK;
          if (tail) {
            const int n_tile = kTileM * tail_row_bytes;
            if (leader) mbar_arrive_expect_tx(full, 2 * (2 * n_tile + half_rows * tail_row_bytes));
            if (quad) {
              if (qrank < 2) tma_load_nd_2sm_mc(p.a_rank, a_dst, &p.tmA_tail, full, c0, twin_mask);
              else tma_load_nd_2sm_mc(p.a_rank, a_dst + n_tile, &p.tmA_tail, full, c1, twin_mask);
            } else {
              tma_load_nd_2sm(p.a_rank, a_dst, &p.tmA_tail, full, c0);
              tma_load_nd_2sm(p.a_rank, a_dst + n_tile, &p.tmA_tail, full, c1);
            }
            tma_load_2d_2sm(a_dst + 2 * n_tile, &p.tmB_tail, full, kc * kBlockK, brow);
            if (++s == p.stages) { s = 0; par ^= 1; }
            continue;
          }
          // the leader arms its barrier with the bytes BOTH CTAs will deliver for this stage
          if (leader) mbar_arrive_expect_tx(full, 2 * (a_bytes + b_bytes + (with_tail ? tail_area : 0)));
          if (quad) {
            if (qrank < 2) tma_load_nd_2sm_mc(p.a_rank, a_dst, &p.tmA, full, c0, twin_mask);
            else tma_load_nd_2sm_mc(p.a_rank, a_dst + a_tile, &p.tmA, full, c1, twin_mask);
          } else {
            tma_load_nd_2sm(p.a_rank, a_dst, &p.tmA, full, c0);
            tma_load_nd_2sm(p.a_rank, a_dst + a_tile, &p.tmA, full, c1);
          }
          tma_load_2d_2sm(a_dst + 2 * a_tile, &p.tmB, full, kc * kBlockK, brow);
          if (with_tail) {
            const uint32_t t_dst = a_dst + a_bytes + b_bytes;
            c0[0] = (kc + 1) * kBlockK;
            c1[0] = (kc + 1) * kBlockK;
            if (quad) {
              if (qrank < 2) tma_load_nd_2sm_mc(p.a_rank, t_dst, &p.tmA_tail, full, c0, twin_mask);
              else tma_load_nd_2sm_mc(p.a_rank, t_dst + t_tile, &p.tmA_tail, full, c1, twin_mask);
            } else {
              tma_load_nd_2sm(p.a_rank, t_dst, &p.tmA_tail, full, c0);
              tma_load_nd_2sm(p.a_rank, t_dst + t_tile, &p.tmA_tail, full, c1);
            }
            tma_load_2d_2sm(t_dst + 2 * t_tile, &p.tmB_tail, full, (kc + 1) * kBlockK, brow);
          }
          if (++s == p.stages) { s = 0; par ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (leader && elect_one()) {
      // ------------------------------------------------------------------ MMA issuer (leader CTA only)
      const uint32_t idesc = make_idesc_bf16(2 * kTileM, p.bn_tile, 0, 0);
      const int tail_steps = (p.k_total - (p.kchunks - 1) * kBlockK + 15) / 16;
      const uint32_t smem0 = smem_u32(smem);
      const uint64_t adesc0 = make_smem_desc_sw128(smem0, 16, 1024);
      const uint64_t bdesc0 = make_smem_desc_sw128(smem0 + a_bytes, 16, 1024);
      const uint32_t desc_step = (uint32_t)stage_bytes >> 4;
      const uint32_t full0 = smem_u32(&ps->full[0]), empty0 = smem_u32(&ps->empty[0]);
      int s = 0, kc = 0;
      uint32_t par = 0, acc = 0;
      for (int it = 0; it < iters; ++it) {
        mbar_wait(full0 + 8 * s, par);
        tc_fence_after();
        if (tr && it == 0) ps->trace[1] = clock64();
        const uint64_t adesc = adesc0 + (uint64_t)(desc_step * s);
        const uint64_t bdesc = bdesc0 + (uint64_t)(desc_step * s);
        if (!p.merge_tail && p.tail_mode != 0 && kc == p.kchunks - 1) {
          // narrow tail tile: rows of 32 B (SWIZZLE_32B, 8-row groups 256 B apart) or 64 B (SWIZZLE_64B, 512 B)
          const uint32_t base = smem0 + s * stage_bytes;
          const uint32_t n_tile = kTileM * tail_row_bytes;
          const uint32_t layout = p.tail_mode == 1 ? 6u : 4u;
          const uint32_t sbo = 8 * tail_row_bytes;
          const uint64_t ta0 = make_smem_desc(base, 16, sbo, layout);
          const uint64_t ta1 = make_smem_desc(base + n_tile, 16, sbo, layout);
          const uint64_t tb = make_smem_desc(base + 2 * n_tile, 16, sbo, layout);
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            if (k < tail_steps) {
              umma_bf16_2sm(tmem, ta0 + 2 * k, tb + 2 * k, idesc, acc);
              umma_bf16_2sm(tmem + kTmemCols, ta1 + 2 * k, tb + 2 * k, idesc, acc);
              acc = 1;
            }
          }
          umma_commit_2sm(empty0 + 8 * s, quad ? (uint16_t)0xF : pair_mask);
          if (++kc == kloops) kc = 0;
          if (++s == p.stages) { s = 0; par ^= 1; }
          continue;
        }
        const int nsteps = (p.merge_tail || kc != p.kchunks - 1 || p.tail_mode == 0) ? kBlockK / 16 : tail_steps;
#pragma unroll
        for (int k = 0; k < kBlockK / 16; ++k) {
          if (k < nsteps) {
            // pair tile 0 -> TMEM columns [0, N), pair tile 1 (A + 16 KiB) -> [256, 256 + N), in both CTAs
            umma_bf16_2sm(tmem, adesc + 2 * k, bdesc + 2 * k, idesc, acc);
            umma_bf16_2sm(tmem + kTmemCols, adesc + (kABytes >> 4) + 2 * k, bdesc + 2 * k, idesc, acc);
            acc = 1;
          }
        }
        if (p.merge_tail && kc == kloops - 1) {
          // the tap's 16-wide tail (SWIZZLE_32B tiles behind the full tiles of this stage): one K step
          const uint32_t tb0 = smem0 + s * stage_bytes + a_bytes + b_bytes;
          const uint32_t t_tile = kTileM * 32;
          const uint64_t tbd = make_smem_desc(tb0 + 2 * t_tile, 16, 256, 6);
          umma_bf16_2sm(tmem, make_smem_desc(tb0, 16, 256, 6), tbd, idesc, 1);
          umma_bf16_2sm(tmem + kTmemCols, make_smem_desc(tb0 + t_tile, 16, 256, 6), tbd, idesc, 1);
        }
        umma_commit_2sm(empty0 + 8 * s, quad ? (uint16_t)0xF : pair_mask);   // stage free (in all CTAs that fill it)
        if (++kc == kloops) kc = 0;
        if (++s == p.stages) { s = 0; par ^= 1; }
      }
      if (tr) ps->trace[2] = clock64();
      umma_commit_2sm(smem_u32(&ps->tmem_full), pair_mask);   // accumulators final in both CTAs of the pair
    }
    __syncwarp();
  } else {
    // -------------------------------------------------------------------- epilogue (each CTA: its own rows)
    const int q = warp & 3;
    const int cg = (warp - 2) >> 2;
    const int ncg = ((int)(blockDim.x >> 5) - 2) >> 2;
    const int r = q * 32 + lane;
    const int iw = r % p.bw;
    const int ih = (r / p.bw) % p.bh;
    const int in = r / (p.bw * p.bh);
    EpilogueArgs ea;
    ea.bias = p.bias; ea.act = p.act; ea.leak = p.leak; ea.mask_src = p.mask_src;
    ea.mask_kind = p.mask_kind; ea.alpha = p.alpha; ea.out = p.out; ea.out_f32 = p.out_f32;
    ea.accumulate = p.accumulate; ea.ncols = p.ncols; ea.pipelined = p.epi_pipe;
    ea.mask_bits = p.mask_bits; ea.bits_out = p.bits_out; ea.bits_pitch = p.bits_pitch; ea.row_elems = p.row_elems;
    ea.stage_row = 0; ea.stage_bits = 0; ea.stage_col0 = 0;
    ea.bias_smem = 0; ea.bias_col0 = n0;
    ea.partial_row = nullptr; ea.partial_n = 0; ea.partial_stride = 0;
    if (p.bias) {
      for (int i = (int)threadIdx.x - 64; i < p.bn_tile; i += (int)blockDim.x - 64)
        ps->bias[i] = (n0 + i < p.ncols) ? __ldg(p.bias + n0 + i) : 0.f;
      named_barrier(2, (int)blockDim.x - 64);
      ea.bias_smem = smem_u32(&ps->bias[0]);
    }
    const int total_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
    bool row_ok[2];
    long long off[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int pw = pw0[i] + iw, ph = ph0[i] + ih, pn = pn0[i] + in;
      const bool tile_ok = (int)(blockIdx.x * 2 + i) < total_tiles;
      row_ok[i] = tile_ok && pw < ext_w && ph < ext_h && pn < p.ext_n;
      off[i] = p.phase_o_off[phase] + (long long)pn * p.o_sn + (long long)ph * p.o_sh + (long long)pw * p.o_sw;
      if (row_ok[i] && cg == 0) epilogue_prefetch_mask(ea, off[i], n0, p.bn_tile);
    }
    if (p.b_prefetch && blockIdx.x < 2) {
      const int row0 = n0 + (int)crank * half_rows;
      const int lines = (p.k_total * 2 + 127) >> 7;
      const int per_tap = half_rows * lines;
      const int total = ntaps * per_tap;
      const char* base = reinterpret_cast<const char*>(p.b_base);
      for (int i = (int)threadIdx.x - 64; i < total; i += (int)blockDim.x - 64) {
        const int tp = i / per_tap, rem = i - tp * per_tap;
        const int rr = rem / lines, ln = rem - rr * lines;
        const int row = p.tap_b_row[tap_begin + tp] + row0 + rr;
        if (row < p.b_rows_total)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(base + (size_t)row * p.b_pitch_bytes + ln * 128));
      }
    }
    mbar_wait(smem_u32(&ps->tmem_full), 0);
    tc_fence_after();
    if (tr && threadIdx.x == 64) ps->trace[3] = clock64();
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16) + i * kTmemCols;
      if (p.tma_store) {
        ea.stage_row = smem_u32(smem) + (uint32_t)((i * kTileM + r) * p.stage_pitch);
        ea.stage_col0 = n0;
      }
      epilogue_row<kSimple>(ea, trow, off[i], row_ok[i], n0, cg * 16, ncg * 16, min(p.bn_tile, p.ncols - n0));
    }
    if (tr && threadIdx.x == 64) ps->trace[4] = clock64();
    if (p.tma_store) {
      fence_proxy_async_smem();
      named_barrier(1, (int)blockDim.x - 64);
      if (threadIdx.x == 64) {
        for (int i = 0; i < 2; ++i)
          if ((int)(blockIdx.x * 2 + i) < total_tiles)
            tma_store_4d(&p.tmOut[phase], smem_u32(smem) + (uint32_t)(i * kTileM * p.stage_pitch), n0, pw0[i], ph0[i],
                         pn0[i]);
        bulk_commit();
        bulk_wait_read0();
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (tr && threadIdx.x == 0) {
    const long long t0 = ps->trace[0];
    printf("tapgemm2sm grid (%d,%d,%d) iters %d: first_mma %lld  last_mma_issued %lld  epi_start %lld  epi_rows_done %lld  end %lld\n",
           gridDim.x, gridDim.y, gridDim.z, iters, ps->trace[1] - t0, ps->trace[2] - t0, ps->trace[3] - t0,
           ps->trace[4] - t0, (long long)clock64() - t0);
  }
  cluster_sync_all();       // the pair's MMAs / barrier traffic are finished in both CTAs
  if (warp == 1) tmem_dealloc_2sm<2 * kTmemCols>(tmem);
}

// =============================================================================================
// Tap GEMM, 2-CTA form, persistent schedule (whole items, or stream-K pieces)
// =============================================================================================
// The plain 2-CTA kernel launches one cluster per item = (phase, N tile, group of 4 pixel tiles) and every CTA pays
// its own prologue (barrier init, TMEM allocation, cluster sync, pipeline fill) and an epilogue nothing overlaps.
// For items of 8-32 K-iterations (pix2pix's 64..256-channel k4 layers, the autoencoders) that fixed cost is 3-4x
// the MMA time.  Here 74 persistent pairs each walk a contiguous range of the linearised (item, K-iteration) space
// (as wgrad2sm_kernel does) while the TMA ring streams across item boundaries, and
//   * p.sk_snap = 1 (the default use): range boundaries are snapped to item boundaries, so every piece is a whole
//     item; with N tiles <= 128 columns the two pixel tiles of an item take 256 TMEM columns and the accumulators
//     are DOUBLE-BUFFERED: the epilogue of item i overlaps the MMAs of item i+1;
//   * p.sk_snap = 0 (stream-K proper, B200GAN_STREAMK=1): ranges are equal and cut items.  A piece that ends before
//     its item does ("contributor") dumps its fp32 accumulators to a workspace region and raises a flag; the piece
//     with the item's last iteration ("finisher") adds the contributors' partials and runs the fused epilogue.  Only
//     the LAST piece of a range can be a contributor and only the FIRST a finisher with contributors; the contributor
//     piece is processed first, so every flag a finisher waits for was raised long before.  Measured on IWGAN's
//     75-150-iteration items: slower than the plain kernel (profiles/r2_streamk_ab.txt), hence opt-in.
template <bool kSimple>
__global__ void __launch_bounds__(kMaxThreads, 1) tapgemm2sm_sk_kernel(const __grid_constant__ TapGemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();
  const bool leader = crank == 0;
  const int pair = blockIdx.x >> 1;

  const int half_rows = p.bn_tile / 2;
  const int b_bytes = half_rows * kBlockK * 2;
  const int a_bytes = 2 * kABytes;
  const int tail_area = p.merge_tail ? (2 * kTileM + half_rows) * 32 : 0;
  const int stage_bytes = a_bytes + b_bytes + tail_area;
  PipeSmem* ps = reinterpret_cast<PipeSmem*>(smem + (size_t)p.stages * stage_bytes);
  const int kloops = p.merge_tail ? p.kchunks - 1 : p.kchunks;
  const int tail_row_bytes = p.tail_mode == 1 ? 32 : 64;
  // accumulator buffers: N tile <= 128 -> buffer b at columns [256 b, 256 b + 256), tile i at +128 i
  const int nbuf = p.bn_tile <= 128 ? 2 : 1;
  const uint32_t tile_cols = nbuf == 2 ? 128u : (uint32_t)kTmemCols;

  if (threadIdx.x == 0) {
    // this pair's pieces, in processing order (identical in both CTAs)
    auto phase_of = [&](long long pos) {
      int ph = 0;
      while (ph + 1 < p.nphases && pos >= p.sk_phase_base[ph + 1]) ++ph;
      return ph;
    };
    auto len_of = [&](int ph) { return (p.phase_tap_begin[ph + 1] - p.phase_tap_begin[ph]) * kloops; };
    auto snap = [&](long long pos) {            // floor to the start of the item that contains pos
      if (pos >= p.sk_total) return p.sk_total;
      const int ph = phase_of(pos);
      const long long rel = pos - p.sk_phase_base[ph];
      return pos - rel % len_of(ph);
    };
    long long rb = (long long)pair * p.sk_range;
    long long re = min(rb + p.sk_range, p.sk_total);
    if (p.sk_snap) { rb = snap(rb); re = snap(re); }
    int n = 0;
    for (long long pos = rb; pos < re && n < kMaxSkSegs;) {
      const int ph = phase_of(pos);
      const int len_p = len_of(ph);
      const long long rel = pos - p.sk_phase_base[ph];
      const int item = (int)(rel / len_p);
      SkSeg g;
      g.phase = (uint8_t)ph;
      g.ntile = (uint16_t)(item / p.sk_groups);
      g.group = (uint32_t)(item - (int)g.ntile * p.sk_groups);
      g.it0 = (uint16_t)(rel - (long long)item * len_p);
      g.len = (uint16_t)min((long long)(len_p - g.it0), re - pos);
      g.finishing = (g.it0 + g.len == len_p) ? 1 : 0;
      g.first_pair = (uint16_t)((pos - g.it0) / p.sk_range);
      g.pad_[0] = g.pad_[1] = 0;
      ps->segs[n++] = g;
      pos += g.len;
    }
    // a contributor piece (necessarily the last one) is processed first: rotate it to the front
    if (n > 1 && !ps->segs[n - 1].finishing) {
      const SkSeg last = ps->segs[n - 1];
      for (int j = n - 1; j > 0; --j) ps->segs[j] = ps->segs[j - 1];
      ps->segs[0] = last;
    }
    ps->nseg = n;
  }
  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&p.tmA);
    tma_prefetch_desc(&p.tmB);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(smem_u32(&ps->full[s]), 1);
      mbar_init(smem_u32(&ps->empty[s]), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(&ps->acc_full[b]), 1);
      mbar_init(smem_u32(&ps->acc_empty[b]), 2 * ((blockDim.x >> 5) - 2));   // the epilogue warps of BOTH CTAs
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc_2sm<2 * kTmemCols>(smem_u32(&ps->tmem_base));
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = ps->tmem_base;
  const int nseg = ps->nseg;

  // first pixel of the CTA's two tiles of a group, per piece
  auto tile_origin = [&](const SkSeg& g, int i, int& pw0, int& ph0, int& pn0) {
    int t = (int)g.group * 4 + (int)crank * 2 + i;
    const int tw = t % p.tiles_w; t /= p.tiles_w;
    const int th = t % p.tiles_h; t /= p.tiles_h;
    pw0 = tw * p.bw; ph0 = th * p.bh; pn0 = t * p.bn;
  };

  if (warp == 0) {
    if (elect_one()) {
      // ------------------------------------------------------------------ TMA producer (both CTAs)
      int s = 0;
      uint32_t par = 0;
      const uint32_t full0 = smem_u32(&ps->full[0]), empty0 = smem_u32(&ps->empty[0]);
      const uint32_t smem0 = smem_u32(smem);
      const int a_tile = kABytes, t_tile = kTileM * 32;
      for (int k = 0; k < nseg; ++k) {
        const SkSeg g = ps->segs[k];
        int base[2][5];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          int pw0, ph0, pn0;
          tile_origin(g, i, pw0, ph0, pn0);
          base[i][0] = 0;
#pragma unroll
          for (int d = 0; d < 4; ++d)
            base[i][d + 1] = pw0 * p.a_mul[0][d] + ph0 * p.a_mul[1][d] + pn0 * p.a_mul[2][d];
        }
        const int tap_begin = p.phase_tap_begin[g.phase];
        const int n0 = (int)g.ntile * p.bn_tile;
        int tp = (int)g.it0 / kloops, kc = (int)g.it0 - tp * kloops;
        int c0[5], c1[5], brow = 0;
        bool new_tap = true;
        for (int it = 0; it < (int)g.len; ++it) {
          if (new_tap) {
            const int tap = tap_begin + tp;
#pragma unroll
            for (int d = 0; d < 4; ++d) {
              c0[d + 1] = base[0][d + 1] + p.tap_a_off[tap][d];
              c1[d + 1] = base[1][d + 1] + p.tap_a_off[tap][d];
            }
            brow = p.tap_b_row[tap] + n0 + (int)crank * half_rows;
            new_tap = false;
          }
          mbar_wait(empty0 + 8 * s, par ^ 1);
          const uint32_t full = full0 + 8 * s;
          const uint32_t a_dst = smem0 + s * stage_bytes;
          const bool with_tail = p.merge_tail && kc == kloops - 1;
          const bool tail = !p.merge_tail && p.tail_mode != 0 && kc == p.kchunks - 1;
          c0[0] = kc * kBlockK;
          c1[0] = kc * kBlockK;
          if (tail) {
            const int n_tile = kTileM * tail_row_bytes;
            if (leader) mbar_arrive_expect_tx(full, 2 * (2 * n_tile + half_rows * tail_row_bytes));
            tma_load_nd_2sm(p.a_rank, a_dst, &p.tmA_tail, full, c0);
            tma_load_nd_2sm(p.a_rank, a_dst + n_tile, &p.tmA_tail, full, c1);
            tma_load_2d_2sm(a_dst + 2 * n_tile, &p.tmB_tail, full, kc * kBlockK, brow);
          } else {
            if (leader) mbar_arrive_expect_tx(full, 2 * (a_bytes + b_bytes + (with_tail ? tail_area : 0)));
            tma_load_nd_2sm(p.a_rank, a_dst, &p.tmA, full, c0);
            tma_load_nd_2sm(p.a_rank, a_dst + a_tile, &p.tmA, full, c1);
            tma_load_2d_2sm(a_dst + 2 * a_tile, &p.tmB, full, kc * kBlockK, brow);
            if (with_tail) {
              const uint32_t t_dst = a_dst + a_bytes + b_bytes;
              c0[0] = (kc + 1) * kBlockK;
              c1[0] = (kc + 1) * kBlockK;
              tma_load_nd_2sm(p.a_rank, t_dst, &p.tmA_tail, full, c0);
              tma_load_nd_2sm(p.a_rank, t_dst + t_tile, &p.tmA_tail, full, c1);
              tma_load_2d_2sm(t_dst + 2 * t_tile, &p.tmB_tail, full, (kc + 1) * kBlockK, brow);
            }
          }
          if (++s == p.stages) { s = 0; par ^= 1; }
          if (++kc == kloops) { kc = 0; ++tp; new_tap = true; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (leader && elect_one()) {
      // ------------------------------------------------------------------ MMA issuer (leader CTA only)
      const uint32_t idesc = make_idesc_bf16(2 * kTileM, p.bn_tile, 0, 0);
      const int tail_steps = (p.k_total - (p.kchunks - 1) * kBlockK + 15) / 16;
      const uint32_t smem0 = smem_u32(smem);
      const uint64_t adesc0 = make_smem_desc_sw128(smem0, 16, 1024);
      const uint64_t bdesc0 = make_smem_desc_sw128(smem0 + a_bytes, 16, 1024);
      const uint32_t desc_step = (uint32_t)stage_bytes >> 4;
      const uint32_t full0 = smem_u32(&ps->full[0]), empty0 = smem_u32(&ps->empty[0]);
      int s = 0;
      uint32_t par = 0;
      for (int k = 0; k < nseg; ++k) {
        const SkSeg g = ps->segs[k];
        const int buf = k % nbuf;
        if (k >= nbuf) {
          // both CTAs' epilogues have drained this buffer's previous accumulators
          mbar_wait(smem_u32(&ps->acc_empty[buf]), (uint32_t)((k / nbuf - 1) & 1));
          tc_fence_after();
        }
        const uint32_t d0 = tmem + (uint32_t)buf * kTmemCols, d1 = d0 + tile_cols;
        int kc = (int)g.it0 % kloops;
        uint32_t acc = 0;
        for (int it = 0; it < (int)g.len; ++it) {
          mbar_wait(full0 + 8 * s, par);
          tc_fence_after();
          const uint64_t adesc = adesc0 + (uint64_t)(desc_step * s);
          const uint64_t bdesc = bdesc0 + (uint64_t)(desc_step * s);
          if (!p.merge_tail && p.tail_mode != 0 && kc == p.kchunks - 1) {
            const uint32_t base = smem0 + s * stage_bytes;
            const uint32_t n_tile = kTileM * tail_row_bytes;
            const uint32_t layout = p.tail_mode == 1 ? 6u : 4u;
            const uint32_t sbo = 8 * tail_row_bytes;
            const uint64_t ta0 = make_smem_desc(base, 16, sbo, layout);
            const uint64_t ta1 = make_smem_desc(base + n_tile, 16, sbo, layout);
            const uint64_t tb = make_smem_desc(base + 2 * n_tile, 16, sbo, layout);
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              if (q < tail_steps) {
                umma_bf16_2sm(d0, ta0 + 2 * q, tb + 2 * q, idesc, acc);
                umma_bf16_2sm(d1, ta1 + 2 * q, tb + 2 * q, idesc, acc);
                acc = 1;
              }
            }
          } else {
            const int nsteps = (p.merge_tail || kc != p.kchunks - 1 || p.tail_mode == 0) ? kBlockK / 16 : tail_steps;
#pragma unroll
            for (int q = 0; q < kBlockK / 16; ++q) {
              if (q < nsteps) {
                umma_bf16_2sm(d0, adesc + 2 * q, bdesc + 2 * q, idesc, acc);
                umma_bf16_2sm(d1, adesc + (kABytes >> 4) + 2 * q, bdesc + 2 * q, idesc, acc);
                acc = 1;
              }
            }
            if (p.merge_tail && kc == kloops - 1) {
              const uint32_t tb0 = smem0 + s * stage_bytes + a_bytes + b_bytes;
              const uint32_t t_tile = kTileM * 32;
              const uint64_t tbd = make_smem_desc(tb0 + 2 * t_tile, 16, 256, 6);
              umma_bf16_2sm(d0, make_smem_desc(tb0, 16, 256, 6), tbd, idesc, 1);
              umma_bf16_2sm(d1, make_smem_desc(tb0 + t_tile, 16, 256, 6), tbd, idesc, 1);
            }
          }
          umma_commit_2sm(empty0 + 8 * s, (uint16_t)0x3);
          if (++kc == kloops) kc = 0;
          if (++s == p.stages) { s = 0; par ^= 1; }
        }
        umma_commit_2sm(smem_u32(&ps->acc_full[buf]), (uint16_t)0x3);
      }
    }
    __syncwarp();
  } else {
    // -------------------------------------------------------------------- epilogue warps (each CTA: its own rows)
    const int q = warp & 3;
    const int cg = (warp - 2) >> 2;
    const int ncg = ((int)(blockDim.x >> 5) - 2) >> 2;
    const int nepi = (int)blockDim.x - 64;
    const int r = q * 32 + lane;
    const int iw = r % p.bw;
    const int ih = (r / p.bw) % p.bh;
    const int in = r / (p.bw * p.bh);
    const int total_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
    EpilogueArgs ea;
    ea.bias = p.bias; ea.act = p.act; ea.leak = p.leak; ea.mask_src = p.mask_src;
    ea.mask_kind = p.mask_kind; ea.alpha = p.alpha; ea.out = p.out; ea.out_f32 = p.out_f32;
    ea.accumulate = p.accumulate; ea.ncols = p.ncols; ea.pipelined = p.epi_pipe;
    ea.mask_bits = p.mask_bits; ea.bits_out = p.bits_out; ea.bits_pitch = p.bits_pitch; ea.row_elems = p.row_elems;
    ea.stage_row = 0; ea.stage_bits = 0; ea.stage_col0 = 0;
    ea.bias_smem = 0; ea.bias_col0 = 0;
    ea.partial_row = nullptr; ea.partial_n = 0; ea.partial_stride = 2 * p.sk_region;     // next pair, same CTA rank
    float* my_region = p.sk_partial + ((long long)pair * 2 + crank) * p.sk_region;
    int bias_ntile = -1;
    for (int k = 0; k < nseg; ++k) {
      const SkSeg g = ps->segs[k];
      const int buf = k % nbuf;
      const int n0 = (int)g.ntile * p.bn_tile;
      const int ext_w = p.phase_ext_w[g.phase], ext_h = p.phase_ext_h[g.phase];
      bool row_ok[2];
      long long off[2];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        int pw0, ph0, pn0;
        tile_origin(g, i, pw0, ph0, pn0);
        const int pw = pw0 + iw, ph = ph0 + ih, pn = pn0 + in;
        const bool tile_ok = (int)g.group * 4 + (int)crank * 2 + i < total_tiles;
        row_ok[i] = tile_ok && pw < ext_w && ph < ext_h && pn < p.ext_n;
        off[i] = p.phase_o_off[g.phase] + (long long)pn * p.o_sn + (long long)ph * p.o_sh + (long long)pw * p.o_sw;
        if (g.finishing && row_ok[i] && cg == 0) epilogue_prefetch_mask(ea, off[i], n0, p.bn_tile);
      }
      if (g.finishing && p.bias && (int)g.ntile != bias_ntile) {
        named_barrier(2, nepi);                                  // everyone is done with the previous bias row
        for (int i = (int)threadIdx.x - 64; i < p.bn_tile; i += nepi)
          ps->bias[i] = (n0 + i < p.ncols) ? __ldg(p.bias + n0 + i) : 0.f;
        named_barrier(2, nepi);
        bias_ntile = (int)g.ntile;
      }
      ea.bias_smem = p.bias ? smem_u32(&ps->bias[0]) : 0u;
      ea.bias_col0 = n0;
      ea.partial_n = 0;
      if (g.finishing && g.it0 > 0) {
        // wait for the contributors (they ran their piece first: the flags are up unless something is badly off)
        if (threadIdx.x == 64) {
          for (int c = (int)g.first_pair; c < pair; ++c) {
            volatile int* f = p.sk_flags + c * 2 + crank;
            uint32_t spin = 0;
            while (*f == 0) { if (++spin > (1u << 26)) __trap(); }
            *f = 0;                                              // consumed: the next launch starts from zero
          }
          __threadfence();
        }
        named_barrier(3, nepi);
        ea.partial_n = pair - (int)g.first_pair;
      }
      mbar_wait(smem_u32(&ps->acc_full[buf]), (uint32_t)((k / nbuf) & 1));
      tc_fence_after();
      const int c_end = g.finishing ? min(p.bn_tile, p.ncols - n0) : p.bn_tile;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)buf * kTmemCols + (uint32_t)i * tile_cols;
        if (g.finishing) {
          ea.partial_row = p.sk_partial + ((long long)g.first_pair * 2 + crank) * p.sk_region +
                           (long long)(i * kTileM + r) * p.bn_tile;
          epilogue_row<kSimple>(ea, trow, off[i], row_ok[i], n0, cg * 16, ncg * 16, c_end);
        } else {
          epilogue_dump_row(trow, my_region + (long long)(i * kTileM + r) * p.bn_tile, cg * 16, ncg * 16, c_end);
        }
      }
      tc_fence_before();
      if (!g.finishing) {
        __threadfence();                                         // partials visible device-wide ...
        named_barrier(3, nepi);
        if (threadIdx.x == 64) {                                 // ... before the flag goes up
          volatile int* f = p.sk_flags + pair * 2 + crank;
          *f = 1;
          __threadfence();
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(smem_u32(&ps->acc_empty[buf]), 0);    // on the leader's barrier
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc_2sm<2 * kTmemCols>(tmem);
}

static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

// Launch configuration that must be applied once PER DEVICE (cudaFuncSetAttribute is per device; a process may drive
// several): true the first time it is called with `flags` on the current device.
constexpr int kMaxDevices = 64;
static bool first_use_on_device(bool* flags) {
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= kMaxDevices) return true;
  if (flags[dev]) return false;
  flags[dev] = true;
  return true;
}
static int device_sms() {
  static int sms[kMaxDevices] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= kMaxDevices) return 148;
  if (!sms[dev]) {
    cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev);
    if (sms[dev] <= 0) sms[dev] = 148;
  }
  return sms[dev];
}

int gemm_threads() {
  static int v = -1;
  if (v < 0) {
    int w = env_int("B200GAN_EPI_WARPS", 16);
    if (w < 4) w = 4;
    if (w > 16) w = 16;
    w = (w / 4) * 4;
    v = 64 + 32 * w;
  }
  return v;
}

int epilogue_pipelined() {
  static int v = -1;
  if (v < 0) v = env_int("B200GAN_EPI_PIPE", 1);
  return v;
}

int l2_prefetch_distance() {
  static int v = -1;
  if (v < 0) v = env_int("B200GAN_PREFETCH", 0);   // measured: prefetch requests compete with the loads (slower)
  return v;
}

static int g_dual_min_pct = -1;      // B200GAN_DUAL_MIN_PCT / b200_set_tuning("dual_min_pct", v)
void set_dual_min_pct(int v) { g_dual_min_pct = v < 0 ? 0 : v; }
int tapgemm_dual(int m_tiles, int iters) {
  // two pixel tiles per CTA (and with them the 2-CTA kernels: a CTA pair works on 4 pixel tiles) when there are at
  // least 4 pixel tiles; below that single-tile CTAs leave fewer SMs idle (tools/tune_layers.py).
  // dual_min_pct 0: whenever possible (the tests reach the 2-CTA kernels at small batch); > 100: never.
  static int v = -1;
  if (v < 0) v = env_int("B200GAN_DUAL", 2);
  if (g_dual_min_pct < 0) g_dual_min_pct = env_int("B200GAN_DUAL_MIN_PCT", 65);
  // short K loops (image-side GEMMs) gain more from two co-resident CTAs per SM than from sharing B
  if (!(v == 2 && m_tiles >= 2 && iters > 4)) return 1;
  if (g_dual_min_pct == 0) return 2;
  if (g_dual_min_pct > 100) return 1;
  return m_tiles >= 4 ? 2 : 1;
}

static int env_cluster() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("B200GAN_CLUSTER");
    v = e ? atoi(e) : 2;
    if (v != 1 && v != 2) v = 2;
  }
  return v;
}

static int pdl_enabled() {
  static int v = -1;
  if (v < 0) v = env_int("B200GAN_PDL", 0);   // measured neutral (21.17 vs 21.16 ms/step): off by default
  return v;
}

// cluster == 1: plain launch.  All GEMM kernels are launched as programmatic dependents of their predecessor
// (B200GAN_PDL): their prologue (barrier init, TMEM allocation, descriptor prefetch) overlaps its tail.
template <typename Params>
static void launch_clustered(void (*kern)(Params), const Params& p, dim3 grid, size_t smem, int cluster,
                             cudaStream_t stream, int cluster_y = 1) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.gridDim = grid;
  cfg.blockDim = dim3(gemm_threads());
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int n = 0;
  if (cluster > 1 || cluster_y > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = cluster;
    attr[n].val.clusterDim.y = cluster_y;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  if (pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  cudaLaunchKernelEx(&cfg, kern, p);
}

int tapgemm_cluster_size(const TapGemmParams& p) {
  // the B half each CTA loads must be whole 8-row swizzle atoms: bn_tile % 16 == 0 always holds
  return env_cluster();
}

int tapgemm_stage_bytes(int dual, int bn_tile, int merge_tail) {
  return dual * kABytes + bn_tile * kBlockK * 2 + (merge_tail ? (dual * kTileM + bn_tile) * 32 : 0);
}

// 2-CTA form: usable with two pixel tiles per CTA
int tapgemm_2sm(int cluster, int dual, int tail_mode, int merge_tail, int bn_tile) {
  static int v = -1;
  if (v < 0) v = env_int("B200GAN_CTA2", 1);
  (void)tail_mode; (void)merge_tail;
  return v && cluster == 2 && dual == 2 && bn_tile % 16 == 0;
}
int tapgemm_stage_bytes_2sm(int bn_tile, int merge_tail) {
  return 2 * kABytes + (bn_tile / 2) * kBlockK * 2 + (merge_tail ? (2 * kTileM + bn_tile / 2) * 32 : 0);
}

// Library-owned scratch of the stream-K schedule: one region of fp32 partial accumulators per CTA (2 tiles x 128 rows
// x up to 256 columns) and one flag per CTA.  Allocated once per process on first use (never during a stream capture:
// such a launch falls back to the plain 2-CTA kernel); launches that use it are ordered by the stream they share.
struct SkScratch { float* partial; int* flags; int pairs; int device; };
static bool sk_scratch(SkScratch* out, cudaStream_t stream) {
  static SkScratch sc = {nullptr, nullptr, 0, -1};
  int dev = 0;
  cudaGetDevice(&dev);
  if (sc.partial && sc.device == dev) { *out = sc; return true; }
  if (sc.partial) return false;                       // a second device in one process: not supported by this scratch
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(stream, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) { cudaGetLastError(); return false; }
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int pairs = sms / 2;
  float* part = nullptr;
  int* flags = nullptr;
  if (cudaMalloc(&part, (size_t)pairs * 2 * 2 * kTileM * kTmemCols * sizeof(float)) != cudaSuccess ||
      cudaMalloc(&flags, (size_t)pairs * 2 * sizeof(int)) != cudaSuccess) { cudaGetLastError(); return false; }
  cudaMemset(flags, 0, (size_t)pairs * 2 * sizeof(int));
  sc = {part, flags, pairs, dev};
  *out = sc;
  return true;
}

// Decide and launch the persistent form (see tapgemm2sm_sk_kernel).  Returns false when the plain kernel should run.
//   * items of <= 64 K-iterations with N tiles <= 128: whole-item ranges + double-buffered accumulators (default on,
//     B200GAN_PERSIST=0 disables);
//   * B200GAN_STREAMK=1: stream-K proper for everything else that does not fill whole waves (opt-in, see the kernel).
static bool launch_tapgemm_sk(const TapGemmParams& p0, cudaStream_t stream) {
  static int streamk = -1, persist = -1;
  if (streamk < 0) { streamk = env_int("B200GAN_STREAMK", 0); persist = env_int("B200GAN_PERSIST", 1); }
  if (!streamk && !persist) return false;
  const int total_tiles = p0.tiles_w * p0.tiles_h * p0.tiles_n;
  const int groups = (total_tiles + 3) / 4;
  const int ny = (p0.ncols + p0.bn_tile - 1) / p0.bn_tile;
  const int kloops = p0.merge_tail ? p0.kchunks - 1 : p0.kchunks;
  if (groups > 60000 || ny > 60000) return false;
  SkScratch sc;
  TapGemmParams p = p0;
  long long total = 0;
  int min_len = 1 << 30, max_len = 0;
  for (int ph = 0; ph < p.nphases; ++ph) {
    const int len = (p.phase_tap_begin[ph + 1] - p.phase_tap_begin[ph]) * kloops;
    p.sk_phase_base[ph] = total;
    total += (long long)groups * ny * len;
    min_len = len < min_len ? len : min_len;
    max_len = len > max_len ? len : max_len;
  }
  p.sk_phase_base[p.nphases] = total;
  const long long items = (long long)groups * ny * p.nphases;
  if (min_len < 1 || max_len > 60000) return false;
  const int sms = device_sms();
  static int pairs_env = -1;
  if (pairs_env < 0) pairs_env = env_int("B200GAN_SK_PAIRS", 0);
  const int pairs = (pairs_env > 0 && pairs_env < sms / 2) ? pairs_env : sms / 2;
  const long long range = (total + pairs - 1) / pairs;
  const bool whole_items = persist && p.bn_tile <= 128 && max_len <= 64 && items >= pairs + pairs / 2 &&
                           range / min_len + 3 <= kMaxSkSegs;
  if (whole_items) {
    p.sk_snap = 1;
    p.sk_partial = nullptr; p.sk_flags = nullptr;
  } else {
    if (!streamk || min_len < 8 || items < 16) return false;
    if (items % pairs == 0 && p.nphases == 1) return false;          // whole waves already: nothing to gain
    if (range < 24 || range / min_len + 3 > kMaxSkSegs) return false;
    if (!sk_scratch(&sc, stream) || pairs > sc.pairs) return false;
    p.sk_snap = 0;
    p.sk_partial = sc.partial; p.sk_flags = sc.flags;
  }
  p.sk = 1; p.sk_groups = groups; p.sk_ny = ny; p.sk_total = total; p.sk_range = range;
  p.sk_region = 2LL * kTileM * p.bn_tile;
  p.tma_store = 0;
  const size_t smem = (size_t)p.stages * tapgemm_stage_bytes_2sm(p.bn_tile, p.merge_tail) + sizeof(PipeSmem) + 1024;
  static bool configured[kMaxDevices] = {false};
  if (first_use_on_device(configured)) {
    cudaFuncSetAttribute(tapgemm2sm_sk_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(tapgemm2sm_sk_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  }
  const int npairs = (int)((total + range - 1) / range);
  if (epilogue_is_simple(p.act, p.mask_src, p.mask_bits, p.out_f32, p.accumulate))
    launch_clustered(tapgemm2sm_sk_kernel<true>, p, dim3(2 * npairs), smem, 2, stream);
  else
    launch_clustered(tapgemm2sm_sk_kernel<false>, p, dim3(2 * npairs), smem, 2, stream);
  return true;
}

void launch_tapgemm(const TapGemmParams& p, cudaStream_t stream) {
  if (p.cta2 && launch_tapgemm_sk(p, stream)) return;
  if (p.cta2) {
    const size_t smem2 = (size_t)p.stages * tapgemm_stage_bytes_2sm(p.bn_tile, p.merge_tail) + sizeof(PipeSmem) + 1024;
    static bool configured2[kMaxDevices] = {false};
    if (first_use_on_device(configured2)) {
      cudaFuncSetAttribute(tapgemm2sm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      cudaFuncSetAttribute(tapgemm2sm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    }
    const int tiles2 = (p.tiles_w * p.tiles_h * p.tiles_n + 1) / 2;
    const int ny = (p.ncols + p.bn_tile - 1) / p.bn_tile;
    dim3 grid((tiles2 + 1) / 2 * 2, ny, p.nphases);
    TapGemmParams q2 = p;
    static int trace2 = -1, quad_env = -1;
    if (trace2 < 0) trace2 = env_int("B200GAN_GEMM_TRACE", 0);
    if (quad_env < 0) quad_env = env_int("B200GAN_QUAD", 1);
    q2.trace = trace2;
    // Quad clusters only when they do not cost a wave: four-CTA clusters pack worse into the GPCs (33 of them = 132 SMs
    // are co-resident, against 74 pairs = 148 SMs), and the shared A tiles make one wave only ~7 % faster
    q2.quad = 0;
    if (quad_env && ny % 2 == 0) {
      static int max_pairs[kMaxDevices] = {0}, max_quads[kMaxDevices] = {0};
      int dev = 0;
      cudaGetDevice(&dev);
      if (dev >= 0 && dev < kMaxDevices) {
        if (!max_pairs[dev]) {
          auto kern = tapgemm2sm_kernel<true>;
          cudaLaunchConfig_t cfg;
          memset(&cfg, 0, sizeof cfg);
          cfg.gridDim = dim3(8, 2, 1); cfg.blockDim = dim3(gemm_threads()); cfg.dynamicSmemBytes = smem2;
          cudaLaunchAttribute at[1];
          at[0].id = cudaLaunchAttributeClusterDimension;
          at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
          cfg.attrs = at; cfg.numAttrs = 1;
          int n = 0;
          if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n <= 0) { cudaGetLastError(); n = device_sms() / 2; }
          max_pairs[dev] = n;
          at[0].val.clusterDim.y = 2;
          n = 0;
          if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n <= 0) { cudaGetLastError(); n = device_sms() / 5; }
          max_quads[dev] = n;
        }
        const long long pairs_n = (long long)grid.x / 2 * grid.y * grid.z, quads_n = pairs_n / 2;
        const long long waves_p = (pairs_n + max_pairs[dev] - 1) / max_pairs[dev];
        const long long waves_q = (quads_n + max_quads[dev] - 1) / max_quads[dev];
        q2.quad = (waves_q * 93 <= waves_p * 100) ? 1 : 0;
      }
    }
    if (epilogue_is_simple(p.act, p.mask_src, p.mask_bits, p.out_f32, p.accumulate))
      launch_clustered(tapgemm2sm_kernel<true>, q2, grid, smem2, 2, stream, q2.quad ? 2 : 1);
    else
      launch_clustered(tapgemm2sm_kernel<false>, q2, grid, smem2, 2, stream, q2.quad ? 2 : 1);
    return;
  }
  const int stage_bytes = tapgemm_stage_bytes(p.dual, p.bn_tile, p.merge_tail);
  const size_t smem = (size_t)p.stages * stage_bytes + sizeof(PipeSmem) + 1024;
  static bool configured[kMaxDevices] = {false};
  if (first_use_on_device(configured)) {
    cudaFuncSetAttribute(tapgemm_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(tapgemm_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(tapgemm_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(tapgemm_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  }
  const int tiles = (p.tiles_w * p.tiles_h * p.tiles_n + p.dual - 1) / p.dual;
  const int ntile_y = (p.ncols + p.bn_tile - 1) / p.bn_tile;
  const bool simple = epilogue_is_simple(p.act, p.mask_src, p.mask_bits, p.out_f32, p.accumulate);
  if (p.cluster == 2) {
    static int cy_env = -1;
    if (cy_env < 0) cy_env = env_int("B200GAN_CLUSTER_Y", 1);
    TapGemmParams q = p;
    static int trace_env = -1;
    if (trace_env < 0) trace_env = env_int("B200GAN_GEMM_TRACE", 0);
    q.trace = trace_env;
    q.cluster_y = (cy_env == 2 && p.dual == 2 && ntile_y >= 2) ? 2 : 1;
    const int nsplit = p.splits > 1 ? p.splits : 1;
    dim3 grid((tiles + 1) / 2 * 2, (ntile_y + q.cluster_y - 1) / q.cluster_y * q.cluster_y, p.nphases * nsplit);
    if (simple) launch_clustered(tapgemm_kernel<2, true>, q, grid, smem, 2, stream, q.cluster_y);
    else launch_clustered(tapgemm_kernel<2, false>, q, grid, smem, 2, stream, q.cluster_y);
  } else {
    dim3 grid(tiles, ntile_y, p.nphases * (p.splits > 1 ? p.splits : 1));
    if (simple) launch_clustered(tapgemm_kernel<1, true>, p, grid, smem, 1, stream);
    else launch_clustered(tapgemm_kernel<1, false>, p, grid, smem, 1, stream);
  }
}

// =============================================================================================
// Persistent small-K GEMM (image-side layers after im2col / before col2im)
// =============================================================================================
struct SmallKSmem {
  uint64_t b_full;
  uint64_t a_full[8];
  uint64_t a_empty[8];
  uint64_t acc_full[2];
  uint64_t acc_empty[2];
  uint32_t tmem_base;
  long long trace[5][12];      // B200GAN_SMALLK_TRACE: per-tile clock stamps of CTA 0 (debug)
  alignas(16) float bias[256]; // the bias row (epilogue)
};

template <bool kSimple>
__global__ void __launch_bounds__(kMaxThreads, 1) smallk_kernel(const __grid_constant__ SmallKParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  griddep_launch_dependents();
  const int b_bytes = p.bn_tile * kBlockK * 2;            // one K chunk of B
  const int slot_bytes = p.kchunks * kABytes;             // one A tile: 128 rows x full K
  uint8_t* smem_b = smem;
  uint8_t* smem_a = smem + (size_t)p.kchunks * b_bytes;
  uint8_t* smem_out = smem_a + (size_t)p.slots * slot_bytes;                     // staged output tile (tma_store)
  const int out_stage_bytes = p.tma_store ? ((kTileM * p.stage_pitch + 127) & ~127) : 0;
  uint8_t* smem_bits = smem_out + out_stage_bytes;
  const int bits_stage_bytes = p.bits_stage ? ((kTileM * p.bits_pitch * 2 + 127) & ~127) : 0;
  SmallKSmem* ps = reinterpret_cast<SmallKSmem*>(smem_bits + bits_stage_bytes);

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&p.tmA);
    tma_prefetch_desc(&p.tmB);
    mbar_init(smem_u32(&ps->b_full), 1);
    for (int s = 0; s < p.slots; ++s) {
      mbar_init(smem_u32(&ps->a_full[s]), 1);
      mbar_init(smem_u32(&ps->a_empty[s]), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(smem_u32(&ps->acc_full[a]), 1);
      mbar_init(smem_u32(&ps->acc_empty[a]), (blockDim.x >> 5) - 2);   // one arrival per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<2 * kTmemCols>(smem_u32(&ps->tmem_base));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = ps->tmem_base;
  griddep_wait();      // everything above touched only shared memory / TMEM / kernel parameters

  if (warp == 0) {
    if (elect_one()) {
      const uint32_t bfull = smem_u32(&ps->b_full);
      mbar_arrive_expect_tx(bfull, p.kchunks * b_bytes);
      for (int kc = 0; kc < p.kchunks; ++kc)
        tma_load_2d(smem_u32(smem_b + (size_t)kc * b_bytes), &p.tmB, bfull, kc * kBlockK, 0);
      int s = 0;
      uint32_t par = 0;
      // A is streamed once from HBM: the few smem slots do not keep enough requests in flight to cover the
      // DRAM latency, so tiles are pulled into L2 p.l2_ahead tiles before their loads
      auto prefetch_tile = [&](int tile) {
        if (tile < p.num_tiles)
          for (int kc = 0; kc < p.kchunks; ++kc) {
            int c[2] = {kc * kBlockK, tile * kTileM};
            tma_prefetch_nd(2, &p.tmA, c);
          }
      };
      for (int i = 0; i < p.l2_ahead; ++i) prefetch_tile(blockIdx.x + i * gridDim.x);
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        if (p.l2_ahead) prefetch_tile(tile + p.l2_ahead * gridDim.x);
        mbar_wait(smem_u32(&ps->a_empty[s]), par ^ 1);
        if (p.trace && blockIdx.x == 0 && tile / (int)gridDim.x < 12) ps->trace[0][tile / gridDim.x] = clock64();
        const uint32_t full = smem_u32(&ps->a_full[s]);
        mbar_arrive_expect_tx(full, slot_bytes);
        const uint32_t dst = smem_u32(smem_a + (size_t)s * slot_bytes);
        for (int kc = 0; kc < p.kchunks; ++kc)
          tma_load_2d(dst + kc * kABytes, &p.tmA, full, kc * kBlockK, tile * kTileM);
        if (++s == p.slots) { s = 0; par ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc = make_idesc_bf16(kTileM, p.bn_tile, 0, 0);
      const int tail_steps = (p.k_total - (p.kchunks - 1) * kBlockK + 15) / 16;
      const uint64_t bdesc0 = make_smem_desc_sw128(smem_u32(smem_b), 16, 1024);
      const uint64_t adesc0 = make_smem_desc_sw128(smem_u32(smem_a), 16, 1024);
      mbar_wait(smem_u32(&ps->b_full), 0);
      int s = 0, acc = 0;
      uint32_t par = 0, accpar = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        mbar_wait(smem_u32(&ps->acc_empty[acc]), accpar ^ 1);   // epilogue has drained this accumulator
        if (p.trace && blockIdx.x == 0 && tile / (int)gridDim.x < 12) ps->trace[1][tile / gridDim.x] = clock64();
        mbar_wait(smem_u32(&ps->a_full[s]), par);
        if (p.trace && blockIdx.x == 0 && tile / (int)gridDim.x < 12) ps->trace[2][tile / gridDim.x] = clock64();
        tc_fence_after();
        const uint64_t adesc = adesc0 + (uint64_t)(((uint32_t)slot_bytes >> 4) * s);
        uint32_t accum = 0;
        for (int kc = 0; kc < p.kchunks; ++kc) {
          const int nsteps = (kc == p.kchunks - 1) ? tail_steps : kBlockK / 16;
          const uint64_t ad = adesc + (uint64_t)((kABytes >> 4) * kc);
          const uint64_t bd = bdesc0 + (uint64_t)(((uint32_t)b_bytes >> 4) * kc);
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {
            if (k < nsteps) {
              umma_bf16(tmem + acc * kTmemCols, ad + 2 * k, bd + 2 * k, idesc, accum);
              accum = 1;
            }
          }
        }
        umma_commit(smem_u32(&ps->a_empty[s]));
        umma_commit(smem_u32(&ps->acc_full[acc]));
        if (++s == p.slots) { s = 0; par ^= 1; }
        acc ^= 1;
        if (acc == 0) accpar ^= 1;
      }
    }
    __syncwarp();
  } else {
    const int q = warp & 3;
    const int cg = (warp - 2) >> 2;
    const int ncg = ((int)(blockDim.x >> 5) - 2) >> 2;
    const int r = q * 32 + lane;
    EpilogueArgs ea;
    ea.bias = p.bias; ea.act = p.act; ea.leak = p.leak; ea.mask_src = p.mask_src;
    ea.mask_kind = p.mask_kind; ea.alpha = p.alpha; ea.out = p.out; ea.out_f32 = p.out_f32;
    ea.accumulate = p.accumulate; ea.ncols = p.ncols; ea.pipelined = p.epi_pipe;
    ea.mask_bits = p.mask_bits; ea.bits_out = p.bits_out; ea.bits_pitch = p.bits_pitch; ea.row_elems = p.row_elems;
    ea.stage_row = 0; ea.stage_bits = 0; ea.stage_col0 = 0;
    ea.bias_smem = 0; ea.bias_col0 = 0;
    ea.partial_row = nullptr; ea.partial_n = 0; ea.partial_stride = 0;
    if (p.bias) {
      for (int i = (int)threadIdx.x - 64; i < p.bn_tile; i += (int)blockDim.x - 64)
        ps->bias[i] = (i < p.ncols) ? __ldg(p.bias + i) : 0.f;
      named_barrier(2, (int)blockDim.x - 64);
      ea.bias_smem = smem_u32(&ps->bias[0]);
    }
    int acc = 0;
    uint32_t accpar = 0;
    const int nepi = (int)blockDim.x - 64;
    const bool issuer = threadIdx.x == 64;
    if (p.tma_store) {
      ea.stage_row = smem_u32(smem_out) + (uint32_t)(r * p.stage_pitch);
      ea.stage_bits = p.bits_stage ? smem_u32(smem_bits) + (uint32_t)(r * p.bits_pitch * 2) : 0u;
      ea.stage_col0 = 0;
    }
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const long long row = (long long)tile * kTileM + r;
      const bool row_ok = row < p.M;
      const long long off = row * p.ldo;
      if (row_ok && cg == 0) epilogue_prefetch_mask(ea, off, 0, p.bn_tile);
      mbar_wait(smem_u32(&ps->acc_full[acc]), accpar);
      if (p.trace && blockIdx.x == 0 && threadIdx.x == 64 && tile / (int)gridDim.x < 12) ps->trace[3][tile / gridDim.x] = clock64();
      tc_fence_after();
      if (p.tma_store) {
        if (issuer) bulk_wait_read0();          // the previous tile's bulk store has drained the staging buffer
        named_barrier(1, nepi);
      }
      const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16) + acc * kTmemCols;
      // (this persistent kernel serves the image-side GEMMs: the hot one is bias + lrelu + sign bitmap)
      epilogue_row<kSimple, (1 << 6)>(ea, trow, off, row_ok, 0, cg * 16, ncg * 16, min(p.bn_tile, p.ncols));
      tc_fence_before();
      __syncwarp();
      if (p.trace && blockIdx.x == 0 && threadIdx.x == 64 && tile / (int)gridDim.x < 12) ps->trace[4][tile / gridDim.x] = clock64();
      if (lane == 0) mbar_arrive(smem_u32(&ps->acc_empty[acc]));
      if (p.tma_store) {
        fence_proxy_async_smem();               // staged rows (generic proxy) -> visible to the bulk copy engine
        named_barrier(1, nepi);
        if (issuer) {
          tma_store_2d(&p.tmOut, smem_u32(smem_out), 0, tile * kTileM);
          if (p.bits_stage)
            bulk_store_1d(p.bits_out + (size_t)tile * kTileM * p.bits_pitch, smem_u32(smem_bits),
                          (uint32_t)(kTileM * p.bits_pitch * 2));
          bulk_commit();
        }
      }
      acc ^= 1;
      if (acc == 0) accpar ^= 1;
    }
    if (p.tma_store && issuer) bulk_wait0();
  }

  tc_fence_before();
  __syncthreads();
  if (p.trace && blockIdx.x == 0 && threadIdx.x == 0) {
    const long long t0 = ps->trace[0][0];
    for (int i = 0; i < 7; ++i)
      printf("tile %d: load_issue %lld  mma_acc_free %lld  mma_a_full %lld  epi_start %lld  epi_end %lld\n", i,
             ps->trace[0][i] - t0, ps->trace[1][i] - t0, ps->trace[2][i] - t0, ps->trace[3][i] - t0, ps->trace[4][i] - t0);
  }
  if (warp == 1) tmem_dealloc<2 * kTmemCols>(tmem);
}

bool smallk_fits(int kchunks, int bn_tile, int* slots, int stage_bytes) {
  if (getenv("B200GAN_NO_SMALLK")) return false;
  if (kchunks < 1 || kchunks > 4 || bn_tile > 256) return false;
  const int b_total = kchunks * bn_tile * kBlockK * 2;
  const int slot = kchunks * kABytes;
  int n = (227 * 1024 - 3072 - b_total - stage_bytes) / slot;
  if (n < 2) return false;
  *slots = n > 8 ? 8 : n;
  return true;
}

void launch_smallk(const SmallKParams& p, cudaStream_t stream) {
  const size_t stage = (p.tma_store ? ((kTileM * p.stage_pitch + 127) & ~127) : 0) +
                       (p.bits_stage ? ((kTileM * p.bits_pitch * 2 + 127) & ~127) : 0);
  const size_t smem = (size_t)p.kchunks * p.bn_tile * kBlockK * 2 + (size_t)p.slots * p.kchunks * kABytes + stage +
                      sizeof(SmallKSmem) + 1024;
  static bool configured[kMaxDevices] = {false};
  if (first_use_on_device(configured)) {
    cudaFuncSetAttribute(smallk_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(smallk_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  }
  const int sms = device_sms();
  const int grid = p.num_tiles < sms ? p.num_tiles : sms;
  static int ahead = -1;
  if (ahead < 0) ahead = env_int("B200GAN_SMALLK_AHEAD", 0);
  SmallKParams q = p;
  q.l2_ahead = ahead;
  static int trace = -1;
  if (trace < 0) trace = env_int("B200GAN_SMALLK_TRACE", 0);
  q.trace = trace;
  if (epilogue_is_simple(p.act, p.mask_src, p.mask_bits, p.out_f32, p.accumulate))
    launch_clustered(smallk_kernel<true>, q, dim3(grid), smem, 1, stream);
  else
    launch_clustered(smallk_kernel<false>, q, dim3(grid), smem, 1, stream);
}

// =============================================================================================
// Weight gradient (MN-major operands, split-K, fp32 atomics)
// =============================================================================================
// Work decomposition ("stream-K"): the (unit, K chunk) space -- unit = (filter tap, M tile pair, N tile), K chunks
// = 64-pixel slabs -- is linearised and cut into gridDim.x equal ranges, one per CTA (one CTA per SM).  A range
// that crosses a unit boundary is processed as consecutive segments: the accumulators are flushed to the fp32
// gradient with red.global.add after each segment and the pipeline keeps streaming into the next one.  Every SM
// gets the same number of chunks (no wave quantisation) and each unit is reduced by as few CTAs as possible.
__global__ void __launch_bounds__(kMaxThreads, 1) wgrad_kernel(const __grid_constant__ WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  griddep_launch_dependents();

  constexpr int kBox = 64 * 64 * 2;                     // 8 KiB: 64 pixels x 64 channels
  const int a_bytes = p.dual * 2 * kBox;                // p.dual x 128 "M" channels share one B tile
  const int stage_bytes = a_bytes + p.nb_boxes * kBox;
  PipeSmem* ps = reinterpret_cast<PipeSmem*>(smem + (size_t)p.stages * stage_bytes);

  const long long range_begin = (long long)blockIdx.x * p.chunks_per_cta;
  const long long range_end = min(range_begin + p.chunks_per_cta, (long long)p.units * p.total_chunks);
  if (range_begin >= range_end) return;
  const int md_tiles = (p.m_tiles + p.dual - 1) / p.dual;

  // segment starting at linear position pos: its unit, first chunk and length
  struct Seg { int tap, m0, n0, a_boxes, b_boxes, cbeg, len; };
  auto segment = [&](long long pos) {
    Seg g;
    int u = (int)(pos / p.total_chunks);
    g.cbeg = (int)(pos - (long long)u * p.total_chunks);
    g.len = (int)min((long long)(p.total_chunks - g.cbeg), range_end - pos);
    const int nt = u % p.n_tiles; u /= p.n_tiles;
    const int mt = u % md_tiles; u /= md_tiles;
    g.tap = u;
    g.m0 = mt * p.dual * kTileM; g.n0 = nt * p.bn_tile;
    // number of 64-channel boxes that actually hold data (the rest of the tile is never loaded nor stored;
    // stale smem feeds accumulator rows/columns that the epilogue masks off)
    g.a_boxes = min(2 * p.dual, (p.Ca - g.m0 + 63) / 64);
    g.b_boxes = min(p.nb_boxes, (p.Cb - g.n0 + 63) / 64);
    return g;
  };

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&p.tmA);
    tma_prefetch_desc(&p.tmB);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(smem_u32(&ps->full[s]), 1);
      mbar_init(smem_u32(&ps->empty[s]), 1);
    }
    mbar_init(smem_u32(&ps->tmem_full), 1);
    mbar_init(smem_u32(&ps->tmem_empty), (blockDim.x >> 5) - 2);     // one arrival per epilogue warp
    fence_mbar_init();
  }
  if (warp == 1) {
    if (p.dual == 2) tmem_alloc<2 * kTmemCols>(smem_u32(&ps->tmem_base));
    else tmem_alloc<kTmemCols>(smem_u32(&ps->tmem_base));
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = ps->tmem_base;
  griddep_wait();      // everything above touched only shared memory / TMEM / kernel parameters

  if (warp == 0) {
    if (elect_one()) {
      int s = 0;
      uint32_t par = 0;
      long long prod_wait = 0;
      const long long prod_t0 = clock64();
      const uint32_t full0 = smem_u32(&ps->full[0]), empty0 = smem_u32(&ps->empty[0]);
      const uint32_t smem0 = smem_u32(smem);
      for (long long pos = range_begin; pos < range_end;) {
        const Seg g = segment(pos);
        // running (w,h,n) chunk coordinates: one division per segment, increments afterwards
        int jw, jh, jn;
        {
          int ch = g.cbeg;
          jw = ch % p.chunks_w; ch /= p.chunks_w;
          jh = ch % p.chunks_h; ch /= p.chunks_h;
          jn = ch;
        }
        for (int it = 0; it < g.len; ++it) {
          const int pw0 = jw * p.bw, ph0 = jh * p.bh, pn0 = jn * p.bn;
          const long long tw0 = p.trace ? clock64() : 0;
          mbar_wait(empty0 + 8 * s, par ^ 1);
          if (p.trace) prod_wait += clock64() - tw0;
          const uint32_t full = full0 + 8 * s;
          mbar_arrive_expect_tx(full, (g.a_boxes + g.b_boxes) * kBox);
          const uint32_t a_dst = smem0 + s * stage_bytes;
          int c[5];
#pragma unroll
          for (int d = 0; d < 4; ++d)
            c[d + 1] = p.tap_a_off[g.tap][d] + pw0 * p.a_mul[0][d] + ph0 * p.a_mul[1][d] + pn0 * p.a_mul[2][d];
          for (int b = 0; b < g.a_boxes; ++b) {
            c[0] = g.m0 + b * 64;
            tma_load_nd(p.a_rank, a_dst + b * kBox, &p.tmA, full, c);
          }
          int cb[5] = {0, pw0, ph0, pn0, 0};
          for (int b = 0; b < g.b_boxes; ++b) {
            cb[0] = g.n0 + b * 64;
            tma_load_nd(p.b_rank, a_dst + a_bytes + b * kBox, &p.tmB, full, cb);
          }
          if (++s == p.stages) { s = 0; par ^= 1; }
          if (++jw == p.chunks_w) { jw = 0; if (++jh == p.chunks_h) { jh = 0; ++jn; } }
        }
        pos += g.len;
      }
      if (p.trace && blockIdx.x == 0)
        printf("wgrad producer: total %lld cycles, waiting for free stages %lld, chunks %lld\n",
               (long long)clock64() - prod_t0, prod_wait, range_end - range_begin);
    }
    __syncwarp();
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc = make_idesc_bf16(kTileM, p.bn_tile, 1, 1);
      const uint32_t smem0 = smem_u32(smem);
      // MN-major SW128: 64-channel blocks kBox apart (LBO), 8-pixel groups 1024 B apart (SBO)
      const uint64_t adesc0 = make_smem_desc_sw128(smem0, kBox, 1024);
      const uint64_t bdesc0 = make_smem_desc_sw128(smem0 + a_bytes, kBox, 1024);
      const uint32_t desc_step = (uint32_t)stage_bytes >> 4;
      const uint32_t full0 = smem_u32(&ps->full[0]), empty0 = smem_u32(&ps->empty[0]);
      int s = 0;
      uint32_t par = 0, seg_par = 0;
      bool first = true;
      long long mma_wait = 0;
      const long long mma_t0 = clock64();
      for (long long pos = range_begin; pos < range_end;) {
        const Seg g = segment(pos);
        const bool second = g.a_boxes > 2;       // the second 128-channel tile exists
        if (!first) {
          // the epilogue warps have read the previous segment's accumulators out of TMEM
          mbar_wait(smem_u32(&ps->tmem_empty), seg_par);
          tc_fence_after();
        }
        uint32_t acc = 0;
        for (int it = 0; it < g.len; ++it) {
          const long long tw0 = p.trace ? clock64() : 0;
          mbar_wait(full0 + 8 * s, par);
          if (p.trace) mma_wait += clock64() - tw0;
          tc_fence_after();
          const uint64_t adesc = adesc0 + (uint64_t)(desc_step * s);
          const uint64_t bdesc = bdesc0 + (uint64_t)(desc_step * s);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            // 16 pixels per MMA = two 8-pixel groups = 2048 bytes (addr field is >>4); the second channel
            // tile (A + 16 KiB) accumulates into TMEM columns [256, 256 + N)
            umma_bf16(tmem, adesc + 128 * k, bdesc + 128 * k, idesc, acc);
            if (second) umma_bf16(tmem + kTmemCols, adesc + ((2 * kBox) >> 4) + 128 * k, bdesc + 128 * k, idesc, acc);
            acc = 1;
          }
          umma_commit(empty0 + 8 * s);
          if (++s == p.stages) { s = 0; par ^= 1; }
        }
        umma_commit(smem_u32(&ps->tmem_full));
        if (!first) seg_par ^= 1;
        first = false;
        pos += g.len;
      }
      if (p.trace && blockIdx.x == 0)
        printf("wgrad mma: total %lld cycles, waiting for data %lld\n", (long long)clock64() - mma_t0, mma_wait);
    }
    __syncwarp();
  } else {
    const int q = warp & 3;
    const int cg = (warp - 2) >> 2;
    const int ncg = ((int)(blockDim.x >> 5) - 2) >> 2;
    const bool vec_ok = (p.ldo & 3) == 0 && ((reinterpret_cast<uintptr_t>(p.out) & 15) == 0) &&
                        ((p.out_tap_stride & 3) == 0);
    uint32_t full_par = 0;
    for (long long pos = range_begin; pos < range_end;) {
      const Seg g = segment(pos);
      mbar_wait(smem_u32(&ps->tmem_full), full_par);
      full_par ^= 1;
      tc_fence_after();
      for (int i = 0; i < p.dual; ++i) {
        if (g.m0 + i * kTileM >= p.Ca) break;
        const int ca = g.m0 + i * kTileM + q * 32 + lane;
        const bool row_ok = ca < p.Ca;
        float* orow = p.out + (long long)g.tap * p.out_tap_stride + (long long)ca * p.ldo;
        const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16) + i * kTmemCols;
        // chunks cg, cg+ncg, ...: the TMEM load of the next chunk is in flight while this one is reduced
        auto reduce16 = [&](const uint32_t* v, int col) {
          if (!row_ok) return;
          if (vec_ok && col + 16 <= p.Cb) {
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(orow + col + j),
                           "f"(__uint_as_float(v[j]) * p.alpha), "f"(__uint_as_float(v[j + 1]) * p.alpha),
                           "f"(__uint_as_float(v[j + 2]) * p.alpha), "f"(__uint_as_float(v[j + 3]) * p.alpha)
                           : "memory");
            }
          } else {
            for (int j = 0; j < 16 && col + j < p.Cb; ++j) atomicAdd(orow + col + j, __uint_as_float(v[j]) * p.alpha);
          }
        };
        const int c_end = min(p.bn_tile, p.Cb - g.n0), step = ncg * 16;
        int c = cg * 16;
        if (c < c_end) {
          uint32_t va[16], vb[16];
          tmem_ld16(trow + c, va);
          while (true) {
            const int c1 = c + step;
            tmem_ld_wait16(va);
            if (c1 < c_end) tmem_ld16(trow + c1, vb);
            reduce16(va, g.n0 + c);
            if (c1 >= c_end) break;
            const int c2 = c1 + step;
            tmem_ld_wait16(vb);
            if (c2 < c_end) tmem_ld16(trow + c2, va);
            reduce16(vb, g.n0 + c1);
            if (c2 >= c_end) break;
            c = c2;
          }
        }
      }
      // accumulators are in registers / reduced: the MMA warp may overwrite TMEM with the next segment
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&ps->tmem_empty));
      pos += g.len;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    if (p.dual == 2) tmem_dealloc<2 * kTmemCols>(tmem);
    else tmem_dealloc<kTmemCols>(tmem);
  }
}

// =============================================================================================
// Weight gradient, 2-CTA form (tcgen05 cta_group::2), transposed orientation
// =============================================================================================
// M = output channels (dY^T, plain box), N = input channels (X^T, the tap gather), K = pixels.  A cluster pair
// owns a unit of 512 dY channels x n_tile X channels x one filter tap: CTA r holds M tiles {2r, 2r+1} (two
// accumulators in its TMEM) and HALF of the X tile (n_tile/2 channels), so a 64-pixel K chunk costs a CTA
// 32 KB + 16 KB of shared-memory fill for 2 x (256 x n_tile x 64) MACs -- the 1-CTA kernel needs 64 KB for the
// same MMA time and is bound by that fill rate.  Work is cut stream-K style over the 74 pairs; the epilogue
// writes dW[tap][ci][co] with co along the TMEM lanes, i.e. warp-coalesced scalar red.global.add.
__global__ void __launch_bounds__(kMaxThreads, 1) wgrad2sm_kernel(const __grid_constant__ WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();
  const bool leader = crank == 0;

  constexpr int kBox = 64 * 64 * 2;                     // 8 KiB: 64 pixels x 64 channels
  const int half_n = p.bn_tile / 2;                     // X channels held by one CTA
  const int nb_half = (half_n + 63) / 64;               // 64-channel boxes per CTA for its half
  const int a_bytes = 4 * kBox;                         // two 128-channel dY tiles
  const int stage_bytes = a_bytes + nb_half * kBox;
  PipeSmem* ps = reinterpret_cast<PipeSmem*>(smem + (size_t)p.stages * stage_bytes);

  const long long pair = blockIdx.x >> 1;
  const long long range_begin = pair * p.chunks_per_cta;
  const long long range_end = min(range_begin + p.chunks_per_cta, (long long)p.units * p.total_chunks);
  const bool has_work = range_begin < range_end;        // (both CTAs of a pair agree; they must stay for the syncs)
  const int m_units = (p.Cb + 511) / 512;

  struct Seg { int tap, m0, n0, cbeg, len; };
  auto segment = [&](long long pos) {
    Seg g;
    int u = (int)(pos / p.total_chunks);
    g.cbeg = (int)(pos - (long long)u * p.total_chunks);
    g.len = (int)min((long long)(p.total_chunks - g.cbeg), range_end - pos);
    const int nt = u % p.n_tiles; u /= p.n_tiles;
    const int mu = u % m_units; u /= m_units;
    g.tap = u;
    g.m0 = mu * 512; g.n0 = nt * p.bn_tile;
    return g;
  };
  // valid 64-channel boxes of CTA `r` for a segment (identical formulas in both CTAs: the leader arms the
  // barrier with the bytes of both)
  auto a_boxes_of = [&](const Seg& g, int r) { return max(0, min(4, (p.Cb - (g.m0 + r * 256) + 63) / 64)); };
  auto b_boxes_of = [&](const Seg& g, int r) { return max(0, min(nb_half, (p.Ca - (g.n0 + r * half_n) + 63) / 64)); };

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&p.tmA);
    tma_prefetch_desc(&p.tmB);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(smem_u32(&ps->full[s]), 1);
      mbar_init(smem_u32(&ps->empty[s]), 1);
    }
    mbar_init(smem_u32(&ps->tmem_full), 1);
    mbar_init(smem_u32(&ps->tmem_empty), 2 * ((blockDim.x >> 5) - 2));   // the epilogue warps of BOTH CTAs
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc_2sm<2 * kTmemCols>(smem_u32(&ps->tmem_base));
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = ps->tmem_base;

  if (has_work && warp == 0) {
    if (elect_one()) {
      // ------------------------------------------------------------------ TMA producer (both CTAs)
      int s = 0;
      uint32_t par = 0;
      const uint32_t full0 = smem_u32(&ps->full[0]), empty0 = smem_u32(&ps->empty[0]);
      const uint32_t smem0 = smem_u32(smem);
      for (long long pos = range_begin; pos < range_end;) {
        const Seg g = segment(pos);
        const int my_a = a_boxes_of(g, (int)crank), my_b = b_boxes_of(g, (int)crank);
        const int pair_boxes = a_boxes_of(g, 0) + a_boxes_of(g, 1) + b_boxes_of(g, 0) + b_boxes_of(g, 1);
        const int ma = g.m0 + (int)crank * 256, nb = g.n0 + (int)crank * half_n;
        int jw, jh, jn;
        {
          int ch = g.cbeg;
          jw = ch % p.chunks_w; ch /= p.chunks_w;
          jh = ch % p.chunks_h; ch /= p.chunks_h;
          jn = ch;
        }
        for (int it = 0; it < g.len; ++it) {
          const int pw0 = jw * p.bw, ph0 = jh * p.bh, pn0 = jn * p.bn;
          mbar_wait(empty0 + 8 * s, par ^ 1);
          const uint32_t full = full0 + 8 * s;
          if (leader) mbar_arrive_expect_tx(full, pair_boxes * kBox);
          const uint32_t a_dst = smem0 + s * stage_bytes;
          // M side: dY channels [ma, ma + 256) of the chunk's 64 pixels
          int cb[5] = {0, pw0, ph0, pn0, 0};
          for (int b = 0; b < my_a; ++b) {
            cb[0] = ma + b * 64;
            tma_load_nd_2sm(p.b_rank, a_dst + b * kBox, &p.tmB, full, cb);
          }
          // N side: X channels [nb, nb + half_n) of the tap-shifted, strided pixels
          int c[5];
#pragma unroll
          for (int d = 0; d < 4; ++d)
            c[d + 1] = p.tap_a_off[g.tap][d] + pw0 * p.a_mul[0][d] + ph0 * p.a_mul[1][d] + pn0 * p.a_mul[2][d];
          for (int b = 0; b < my_b; ++b) {
            c[0] = nb + b * 64;
            tma_load_nd_2sm(p.a_rank, a_dst + a_bytes + b * kBox, &p.tmA, full, c);
          }
          if (++s == p.stages) { s = 0; par ^= 1; }
          if (++jw == p.chunks_w) { jw = 0; if (++jh == p.chunks_h) { jh = 0; ++jn; } }
        }
        pos += g.len;
      }
    }
    __syncwarp();
  } else if (has_work && warp == 1) {
    if (leader && elect_one()) {
      // ------------------------------------------------------------------ MMA issuer (leader CTA only)
      const uint32_t idesc = make_idesc_bf16(2 * kTileM, p.bn_tile, 1, 1);
      const uint32_t smem0 = smem_u32(smem);
      const uint64_t adesc0 = make_smem_desc_sw128(smem0, kBox, 1024);
      const uint64_t bdesc0 = make_smem_desc_sw128(smem0 + a_bytes, kBox, 1024);
      const uint32_t desc_step = (uint32_t)stage_bytes >> 4;
      const uint32_t full0 = smem_u32(&ps->full[0]), empty0 = smem_u32(&ps->empty[0]);
      int s = 0;
      uint32_t par = 0, seg_par = 0;
      bool first = true;
      for (long long pos = range_begin; pos < range_end;) {
        const Seg g = segment(pos);
        const bool second = p.Cb - g.m0 > 128;       // some CTA has a second 128-channel tile
        if (!first) {
          mbar_wait(smem_u32(&ps->tmem_empty), seg_par);     // both CTAs' epilogues have drained TMEM
          tc_fence_after();
          seg_par ^= 1;
        }
        first = false;
        uint32_t acc = 0;
        for (int it = 0; it < g.len; ++it) {
          mbar_wait(full0 + 8 * s, par);
          tc_fence_after();
          const uint64_t adesc = adesc0 + (uint64_t)(desc_step * s);
          const uint64_t bdesc = bdesc0 + (uint64_t)(desc_step * s);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            umma_bf16_2sm(tmem, adesc + 128 * k, bdesc + 128 * k, idesc, acc);
            if (second) umma_bf16_2sm(tmem + kTmemCols, adesc + ((2 * kBox) >> 4) + 128 * k, bdesc + 128 * k, idesc, acc);
            acc = 1;
          }
          umma_commit_2sm(empty0 + 8 * s, (uint16_t)0x3);
          if (++s == p.stages) { s = 0; par ^= 1; }
        }
        umma_commit_2sm(smem_u32(&ps->tmem_full), (uint16_t)0x3);
        pos += g.len;
      }
    }
    __syncwarp();
  } else if (has_work && warp >= 2) {
    // -------------------------------------------------------------------- epilogue: dW[tap][ci][co] += acc
    const int q = warp & 3;
    const int cg = (warp - 2) >> 2;
    const int ncg = ((int)(blockDim.x >> 5) - 2) >> 2;
    uint32_t full_par = 0;
    for (long long pos = range_begin; pos < range_end;) {
      const Seg g = segment(pos);
      mbar_wait(smem_u32(&ps->tmem_full), full_par);
      full_par ^= 1;
      tc_fence_after();
      const int c_end = min(p.bn_tile, p.Ca - g.n0), step = ncg * 16;
      for (int i = 0; i < 2; ++i) {
        const int co0 = g.m0 + (int)crank * 256 + i * kTileM;
        if (co0 >= p.Cb) break;
        const int co = co0 + q * 32 + lane;
        const bool row_ok = co < p.Cb;
        float* obase = p.out + (long long)g.tap * p.out_tap_stride + co;
        const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16) + i * kTmemCols;
        auto reduce16 = [&](const uint32_t* v, int c) {
          if (!row_ok) return;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int ci = g.n0 + c + j;
            if (ci < p.Ca)
              asm volatile("red.global.add.f32 [%0], %1;" ::"l"(obase + (long long)ci * p.ldo),
                           "f"(__uint_as_float(v[j]) * p.alpha) : "memory");
          }
        };
        int c = cg * 16;
        if (c < c_end) {
          uint32_t va[16], vb[16];
          tmem_ld16(trow + c, va);
          while (true) {
            const int c1 = c + step;
            tmem_ld_wait16(va);
            if (c1 < c_end) tmem_ld16(trow + c1, vb);
            reduce16(va, c);
            if (c1 >= c_end) break;
            const int c2 = c1 + step;
            tmem_ld_wait16(vb);
            if (c2 < c_end) tmem_ld16(trow + c2, va);
            reduce16(vb, c1);
            if (c2 >= c_end) break;
            c = c2;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(smem_u32(&ps->tmem_empty), 0);    // on the leader's barrier
      pos += g.len;
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc_2sm<2 * kTmemCols>(tmem);
}

static int g_wgrad_min_chunks = 0;    // b200_set_tuning("wgrad_min_chunks", v): fewest 64-pixel chunks a CTA of the stream-K cut takes
void set_wgrad_min_chunks(int v) { g_wgrad_min_chunks = v; }
static int wgrad_min_chunks() { return g_wgrad_min_chunks > 0 ? g_wgrad_min_chunks : 4; }
int wgrad_dual(int m_tiles) {
  static int v = -1;
  if (v < 0) v = env_int("B200GAN_DUAL", 2);
  return (v == 2 && m_tiles >= 2) ? 2 : 1;
}

// Chunks per CTA (pair) of the stream-K cut over `units` x `total_chunks` 64-pixel chunks on at most `max_ctas` CTAs.
// A range that crosses a unit boundary costs one more epilogue (a whole fp32 tile of red.global.add: ~7.7 us for
// 2 x 128 x 256, against ~0.6 us per chunk of main loop -- tools/tune_wgrad.py), so besides the plain equal cut
// (every SM busy; right for the long reductions of the IWGAN layers) the cut aligned to the units is costed: every
// unit split into s equal parts (s | total_chunks, units * s CTAs), or whole units per CTA when there are more units
// than CTAs.  pix2pix 4x4x512->1024: 35 -> 12 us, 16x16x512->512: 24 -> 17 us, 32x32x256->512: 28.5 -> 22 us.
static int wgrad_chunks_per_cta(long long units, long long total_chunks, int max_ctas) {
  const long long total = units * total_chunks;
  const double t_chunk = 0.6, t_epi = 7.7;
  long long ctas = total / wgrad_min_chunks();
  if (ctas > max_ctas) ctas = max_ctas;
  if (ctas < 1) ctas = 1;
  const long long cpc_a = (total + ctas - 1) / ctas;
  if (g_wgrad_min_chunks > 0) return (int)cpc_a;                        // forced by the tuning tool
  const bool aligned_a = (cpc_a <= total_chunks) ? (total_chunks % cpc_a == 0) : (cpc_a % total_chunks == 0);
  const long long segs_a = aligned_a ? std::max(1LL, cpc_a / total_chunks) : cpc_a / total_chunks + 2;
  double best = cpc_a * t_chunk + (segs_a - 1) * t_epi;
  long long cpc = cpc_a;
  if (units > max_ctas) {
    const long long upc = (units + max_ctas - 1) / max_ctas;
    const double c = upc * total_chunks * t_chunk + (upc - 1) * t_epi;
    if (c < best) { best = c; cpc = upc * total_chunks; }
  } else {
    for (long long s = 1; s <= total_chunks && units * s <= max_ctas; ++s) {
      if (total_chunks % s) continue;
      const double c = (double)(total_chunks / s) * t_chunk;
      if (c < best) { best = c; cpc = total_chunks / s; }
    }
  }
  return (int)cpc;
}

static void launch_wgrad_2sm(const WgradParams& p0, cudaStream_t stream) {
  WgradParams p = p0;
  static bool configured[kMaxDevices] = {false};
  if (first_use_on_device(configured)) {
    cudaFuncSetAttribute(wgrad2sm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  }
  // transposed orientation: M = Cb (dY channels, units of 512), N = Ca (X channels)
  p.bn_tile = p.bn_tile_t;
  p.n_tiles = (p.Ca + p.bn_tile - 1) / p.bn_tile;
  const int nb_half = (p.bn_tile / 2 + 63) / 64;
  const int stage_bytes = (4 + nb_half) * 64 * 64 * 2;
  p.stages = (227 * 1024 - 3072) / stage_bytes;
  if (p.stages > 8) p.stages = 8;
  const size_t smem = (size_t)p.stages * stage_bytes + sizeof(PipeSmem) + 1024;
  p.units = ((p.Cb + 511) / 512) * p.n_tiles * p.ntaps;
  const long long total = (long long)p.units * p.total_chunks;
  const int sms = device_sms();
  static int pairs_env = -1;        // B200GAN_WGRAD_PAIRS: CTA pairs of the stream-K partition (default: all SMs)
  if (pairs_env < 0) pairs_env = env_int("B200GAN_WGRAD_PAIRS", 0);
  const int max_pairs = (pairs_env > 0 && pairs_env < sms / 2) ? pairs_env : sms / 2;
  p.chunks_per_cta = wgrad_chunks_per_cta(p.units, p.total_chunks, max_pairs);
  const int npairs = (int)((total + p.chunks_per_cta - 1) / p.chunks_per_cta);
  launch_clustered(wgrad2sm_kernel, p, dim3(2 * npairs), smem, 2, stream);
}

void launch_wgrad(const WgradParams& p0, int /*splits_hint*/, cudaStream_t stream) {
  static int w2 = -1;
  if (w2 < 0) w2 = env_int("B200GAN_WGRAD2", 1);
  // the 2-CTA kernel needs at least three 128-channel dY tiles to fill its 512-row unit reasonably
  if (w2 && p0.bn_tile_t > 0 && p0.Cb > 256 && (p0.ldo & 3) == 0) {
    launch_wgrad_2sm(p0, stream);
    return;
  }
  WgradParams p = p0;
  const int stage_bytes = (2 * p.dual + p.nb_boxes) * 64 * 64 * 2;
  const size_t smem = (size_t)p.stages * stage_bytes + sizeof(PipeSmem) + 1024;
  static bool configured[kMaxDevices] = {false};
  if (first_use_on_device(configured)) {
    cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  }
  // equal share of the (unit, chunk) space per CTA, one CTA per SM; a CTA gets at least 4 chunks
  p.units = ((p.m_tiles + p.dual - 1) / p.dual) * p.n_tiles * p.ntaps;
  const long long total = (long long)p.units * p.total_chunks;
  p.chunks_per_cta = wgrad_chunks_per_cta(p.units, p.total_chunks, device_sms());
  static int trace_env = -1;
  if (trace_env < 0) trace_env = env_int("B200GAN_GEMM_TRACE", 0);
  p.trace = trace_env;
  const int grid = (int)((total + p.chunks_per_cta - 1) / p.chunks_per_cta);
  launch_clustered(wgrad_kernel, p, dim3(grid), smem, 1, stream);
}

}  // namespace b200
