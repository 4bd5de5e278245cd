// tcgen05 / TMEM / TMA scan kernel: bank[N,512] bf16  x  queries^T  ->  fused per-class selection.
//
// One persistent CTA (cta_group::1) or CTA pair (cta_group::2, 256 bank rows per MMA) per SM.
//   warp 0   TMA producer : query block once (resident for the whole kernel), then the bank, each
//                           byte exactly once per Q block, 128 rows x 64 k (16 KB, SWIZZLE_128B) per stage
//   warp 1   MMA issuer   : converged warp, one elected lane issues tcgen05.mma kind::f16 (bf16 in, fp32 accumulate in TMEM),
//                           M = 128*ctas, N = padded query columns (<=256), K = 16 per instruction
//   warp 2   TMEM allocator (512 columns = two accumulator buffers of <=256 columns)
//   warp 3   threshold refresher: class thresholds from the score histograms, off the critical path
//   warps 4-11 epilogue   : two warps per TMEM lane quadrant, interleaved 32-column chunks,
//                           double-buffered tcgen05.ld -> process_chunk (epilogue.cuh); the accumulator
//                           buffer is released as soon as a warp's last chunk is in registers
// Pipelines: smem full/empty ring (TMA <-> MMA), TMEM full/empty double buffer (MMA <-> epilogue).
//
// fp32 banks (F32 = true; the reference's own dtype, utils/extras.py:163): the bank is streamed as fp32 (2 KB/row, TMA
// boxes of 128 rows x 32 k into a second ring), warps 8-11 convert every staged box to bf16 (RNE) and write it into
// the K-major SWIZZLE_128B operand ring the MMAs read; the epilogue has twice the time per row and runs on 4 warps.
// The scores are those of bf16-rounded rows (|delta| <= swat_queries::eps_conv, proven in api.cu); the pipeline
// re-scores every candidate exactly in fp32 afterwards (select.cu).
//
// Replaces caption_embeddings.cuda() @ class_prompt.t() + sorted() + walk,
// /root/reference/retrieval/sample_retrieval.py:400, :754, :439-482 for every class at once.
#include <cuda.h>
#include "common.cuh"
#include "epilogue.cuh"
#include "scan_tc.h"
#include "tc_ptx.cuh"

namespace swat {
namespace {

constexpr int kStageBytes = 128 * 128;   // 128 bank rows x 64 bf16
constexpr int kThreads = 384;             // 4 control warps + 8 epilogue warps, or + 4 epilogue + 4 converter warps (F32)
constexpr int kTailBytes = 5120;         // barriers + tables

// ---------------------------------------------------------------------------------------- kernel
template <int kCtas, int RED, bool PART, bool DENSE, bool F32>
__global__ void __launch_bounds__(kThreads, 1)
scan_tc_kernel(const __grid_constant__ CUtensorMap tm_bank, const __grid_constant__ CUtensorMap tm_q, const TcArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned long long* trace = p.trace ? p.trace + static_cast<size_t>(blockIdx.x) * 8 : nullptr;
  if (trace && threadIdx.x == 0) trace[0] = globaltimer_ns();                      // CTA entry
  const bool probing = p.clock_probe != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
  long long probe_c0 = 0;
  uint64_t probe_t0 = 0;
  if (probing) { probe_c0 = clock64(); probe_t0 = globaltimer_ns(); }

  constexpr int kEpiWarps = F32 ? 4 : 8;   // bf16: two per TMEM lane quadrant on interleaved 32-column chunks
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = (kCtas == 2) ? cluster_ctarank() : 0u;
  const int pair_id = blockIdx.x / kCtas, n_pairs = gridDim.x / kCtas;
  // Work decomposition.  Legacy (n_ranges == 0): Q block = pair % n_qb, tiles strided by the pairs of that block.
  // Unit plan (n_ranges > 0, several launches): unit u = unit_base + pair covers Q block u % n_qb and the tiles
  // r, r + R, r + 2R, ... of range r = u / n_qb.  The n_qb pairs of one range run in the same launch and walk the
  // same tiles in step (one HBM read, the rest L2 hits), and every pair gets the same number of equal units.
  // Legacy launches cover the Q blocks [qb_base, qb_base + qb_count): more blocks than pairs take several launches.
  const int unit = p.unit_base + pair_id;
  const int qb = (p.n_ranges > 0) ? unit % p.n_qb : p.qb_base + pair_id % p.qb_count;
  const int pair_in_qb = (p.n_ranges > 0) ? unit / p.n_qb : pair_id / p.qb_count;
  const int pairs_qb = (p.n_ranges > 0) ? p.n_ranges : (n_pairs - pair_id % p.qb_count + p.qb_count - 1) / p.qb_count;
  const int NB = p.n_blk, NBC = NB / kCtas;
  const uint32_t b_chunk_bytes = static_cast<uint32_t>(NBC) * 128u;
  constexpr int kTileRows = 128 * kCtas;
  const int64_t n_tiles = (p.s.n_rows + kTileRows - 1) / kTileRows;
  const int64_t t_first = (pair_in_qb < pairs_qb) ? pair_in_qb : n_tiles;
  // Tile of this pair's iteration `it` (n_tiles = no more).  Static plan: t_first + it * pairs_qb.  Dynamic plan
  // (p.tile_sched; one Q block): the pairs finish 3-5 % apart under a static split (SM position, survivor bursts), so
  // after its first three tiles a pair takes the next unclaimed tile from a global counter.  The leader CTA's producer
  // claims ahead and publishes (iteration, tile) in a small ring in global memory two iterations early; every other warp of the
  // pair peeks at the entry of iteration it + 1 while it works on iteration it, so nobody waits on the L2 round trip.
  const bool dyn = !DENSE && p.tile_sched != nullptr;
  unsigned long long* tile_ring = dyn ? p.tile_sched + 2 + static_cast<size_t>(pair_id) * kTileRing : nullptr;
  auto tile_peek = [&](uint32_t it) -> unsigned long long {
    return (dyn && it > 0) ? ld_relaxed_gpu_u64(tile_ring + (it & (kTileRing - 1))) : 0ull;
  };
  auto tile_get = [&](uint32_t it, unsigned long long peeked) -> int64_t {      // warp-uniform
    if (!dyn) {
      const int64_t t = t_first + static_cast<int64_t>(it) * pairs_qb;
      return t < n_tiles ? t : n_tiles;
    }
    if (it == 0) return t_first;
    uint32_t spins = 0;
    while (static_cast<uint32_t>(peeked >> 32) != it + 1u) {
      if (++spins > 4000000u) __trap();             // a hang becomes a launch failure after a few seconds
      __nanosleep(40);
      peeked = ld_relaxed_gpu_u64(tile_ring + (it & (kTileRing - 1)));
    }
    const int64_t t = static_cast<int64_t>(static_cast<uint32_t>(peeked));
    return t < n_tiles ? t : n_tiles;
  };
  // CTAs serving this Q block in this launch (the refresher warps split the block's classes among them)
  int qb_peers = pairs_qb, qb_rank = pair_in_qb;
  if (p.n_ranges > 0) {
    const int first_u = p.unit_base + ((qb - p.unit_base) % p.n_qb + p.n_qb) % p.n_qb;
    qb_peers = (p.unit_base + n_pairs - 1 - first_u) / p.n_qb + 1;
    qb_rank = (unit - first_u) / p.n_qb;
  }

  // Pairs that walk the same tiles (one per Q block of a tile range) keep in step so that a tile is read from HBM once
  // and served to the others from L2: launch-local ids [grp_first, grp_first + grp_size) share this pair's tiles.
  int grp_first = pair_id, grp_size = 1;
  if (p.progress != nullptr) {
    if (p.n_ranges > 0) {
      const int lo = max(pair_in_qb * p.n_qb, p.unit_base), hi = min((pair_in_qb + 1) * p.n_qb, p.unit_base + n_pairs);
      grp_first = lo - p.unit_base; grp_size = hi - lo;
    } else {
      grp_first = pair_in_qb * p.qb_count; grp_size = min(p.qb_count, n_pairs - grp_first);
    }
  }

  uint8_t* sB = smem;
  uint8_t* sA = smem + p.smem_b_bytes;                        // operand ring (bf16, what the MMAs read)
  uint8_t* sF = sA + p.n_stages * kStageBytes;                // F32: ring of staged fp32 boxes
  uint8_t* tail = sF + (F32 ? p.n_fstages * kStageBytes : 0);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);     // [8]
  uint64_t* empty_bar = full_bar + 8;                         // [8]
  uint64_t* tfull_bar = empty_bar + 8;                        // [2]
  uint64_t* tempty_bar = tfull_bar + 2;                       // [2]
  uint64_t* q_bar = tempty_bar + 2;                           // [1]
  uint64_t* ffull_bar = q_bar + 1;                            // [8] F32
  uint64_t* fempty_bar = ffull_bar + 8;                       // [8] F32
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(fempty_bar + 8);
  uint32_t* s_done = tmem_slot + 1;                           // epilogue warps that finished
  float* s_tau = reinterpret_cast<float*>(tail + 512);        // [256] (+256 spare)
  int32_t* s_cls = reinterpret_cast<int32_t*>(s_tau + 512);   // [256]
  float* s_cnt = reinterpret_cast<float*>(s_cls + 256);       // [256]
  uint32_t* s_end = reinterpret_cast<uint32_t*>(s_cnt + 256); // [8]

  if (threadIdx.x == 0) {
    prefetch_tmap(&tm_bank);
    prefetch_tmap(&tm_q);
    // operand stage full: one TMA transaction (bf16) or, F32, the converter warps of both CTAs (4 warps x 2 half-boxes each)
    for (int i = 0; i < 8; ++i) { mbar_init(smem_u32(&full_bar[i]), F32 ? 8 * kCtas : 1); mbar_init(smem_u32(&empty_bar[i]), 1); }
    for (int i = 0; i < 8; ++i) { mbar_init(smem_u32(&ffull_bar[i]), 1); mbar_init(smem_u32(&fempty_bar[i]), 4); }
    for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&tfull_bar[i]), 1); mbar_init(smem_u32(&tempty_bar[i]), kEpiWarps * kCtas); }
    mbar_init(smem_u32(q_bar), 1);
    *s_done = 0;
    fence_barrier_init();
  }
  for (int c = threadIdx.x; c < 256; c += kThreads) {
    const bool in = c < NB;
    s_cls[c] = in ? p.s.col_class[qb * NB + c] : -1;
    s_cnt[c] = in ? p.s.col_count[qb * NB + c] : 0.0f;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 256; c += kThreads) {
    // thresholds by column: NaN at padding / non-closing columns, the class threshold elsewhere.
    // Every entry is always SOME valid threshold, so the epilogue warps refresh and read this table
    // without any barrier (a stale value is only more conservative).
    // NaN, not +inf, at padding / non-closing columns: a trailing partial chunk reads uninitialised TMEM
    // columns and +inf >= +inf would pass, whereas every comparison with NaN is false.
    float t = PART ? __int_as_float(0x7fc00000) : INFINITY;   // sign-test path: +inf fails (values are finite); compare path: NaN
    if (!DENSE && s_cls[c] >= 0 && s_cnt[c] > 0.0f) t = fast_tau<RED>(f32_dec(ld_cg_u32(&p.s.st.tau_enc[s_cls[c]])), s_cnt[c]);
    s_tau[c] = t;
  }
  if (threadIdx.x < 8) {
    uint32_t m = 0;
    for (int j = 0; j < 32; ++j) if (s_cnt[threadIdx.x * 32 + j] > 0.0f) m |= 1u << j;
    s_end[threadIdx.x] = m;
  }
  if (warp == 2) tmem_alloc<kCtas>(smem_u32(tmem_slot), 512);
  tc_fence_before();
  if (kCtas == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (trace && threadIdx.x == 0) trace[1] = globaltimer_ns();                      // prologue done (tables, TMEM, cluster sync)

  if (warp == 0) {
    // ================================================================== TMA producer
    // The whole warp walks the ring (converged, so addresses stay in uniform registers); one elected lane issues.
    {
      const uint32_t q_bar_a = smem_u32(q_bar);
      const uint32_t q_bar_lead = (kCtas == 2) ? mapa_rank0(q_bar_a) : q_bar_a;
      if (elect_one()) {
        if (rank == 0) mbar_arrive_expect_tx(q_bar_a, 8u * b_chunk_bytes * kCtas);
        for (int kc = 0; kc < 8; ++kc)
          tma_load_2d<kCtas>(smem_u32(sB + kc * b_chunk_bytes), &tm_q, q_bar_lead, kc * 64, qb * NB + static_cast<int>(rank) * NBC,
                             0x14F0000000000000ull /* evict_last: every CTA re-reads the query block */);
      }
      // lockstep: before requesting tile `it`, every pair of the group should have requested tile it - window.  The
      // peers' progress words are fetched one check AHEAD (the load issued at tile it is consumed at tile it + 2), so the
      // L2 round trip never sits in the producer's critical path -- a blocking read here cost 5-25 % of the scan; only a
      // pair that really runs ahead of its group spins (bounded: the window is a hint, never a dependency).
      const bool lockstep = p.progress != nullptr && grp_size > 1 && rank == 0;
      uint32_t peers_seen = 0xffffffffu;           // min over the group, as of the previous check
      auto keep_in_step = [&](uint32_t it) {
        if (!lockstep) return;
        if (lane == 0) st_relaxed_gpu_u32(p.progress + pair_id, it);
        if (it & 1u) return;
        if (it >= static_cast<uint32_t>(p.lock_window) + 2u) {
          const uint32_t need = it - static_cast<uint32_t>(p.lock_window) - 2u;     // the snapshot is two tiles old
          uint32_t seen = __reduce_min_sync(0xffffffffu, peers_seen);
          for (int spin = 0; seen < need && spin < 48; ++spin) {
            __nanosleep(100);
            seen = __reduce_min_sync(0xffffffffu, lane < grp_size ? ld_relaxed_gpu_u32(p.progress + grp_first + lane) : 0xffffffffu);
          }
        }
        peers_seen = lane < grp_size ? ld_relaxed_gpu_u32(p.progress + grp_first + lane) : 0xffffffffu;   // for the next check
      };
      uint32_t tile_it = 0;
      // Dynamic plan, leader CTA: the pair's first three tiles are static (pair, pair + P, pair + 2P); from then on the
      // tiles of iterations it and it + 1 sit in registers and the claim for it + 2 is in flight -- its value is first
      // touched one iteration after the atomic was issued, so the producer never waits on the L2 round trip.  Other
      // CTA: reads the ring.
      const bool claimer = dyn && rank == 0;
      uint32_t tq1 = 0xffffffffu, tq2 = 0xffffffffu;       // tiles of it + 1, it + 2 relative to the running iteration (0xffffffff = none)
      uint32_t raw_claim = 0;                               // lane 0: result of the pending atomic
      bool claim_pending = false;
      unsigned int* tile_counter = reinterpret_cast<unsigned int*>(p.tile_sched);
      auto static_tile = [&](int j) -> uint32_t {
        const int64_t t = t_first + static_cast<int64_t>(j) * n_pairs;
        return t < n_tiles ? static_cast<uint32_t>(t) : 0xffffffffu;
      };
      auto claim_issue = [&]() {
        if (lane == 0) raw_claim = atomicAdd(tile_counter, 1u);
        claim_pending = true;
      };
      auto claim_take = [&]() -> uint32_t {                                  // every lane gets the same tile
        if (!claim_pending) return 0xffffffffu;
        claim_pending = false;
        const uint32_t v = __shfl_sync(0xffffffffu, raw_claim, 0);
        const uint64_t t = static_cast<uint64_t>(v) + 3ull * static_cast<uint64_t>(n_pairs);
        return t < static_cast<uint64_t>(n_tiles) ? static_cast<uint32_t>(t) : 0xffffffffu;
      };
      auto publish = [&](uint32_t it, uint32_t tile) {
        if (lane == 0) st_relaxed_gpu_u64(tile_ring + (it & (kTileRing - 1)), (static_cast<unsigned long long>(it + 1u) << 32) | tile);
      };
      unsigned long long peeked = 0;
      auto first_tile = [&]() -> int64_t {
        if (claimer) {
          tq1 = static_tile(1); tq2 = static_tile(2);
          publish(1, tq1); publish(2, tq2);
          if (tq2 != 0xffffffffu) claim_issue();                             // for iteration 3
        } else {
          peeked = tile_peek(1);
        }
        return t_first;
      };
      auto next_tile = [&](uint32_t it) -> int64_t {                          // it = the iteration that starts now (>= 1)
        if (!dyn) return tile_get(it, 0ull);
        if (claimer) {
          const uint32_t cur = tq1;
          const uint32_t nxt = claim_take();                                // tile of it + 2, claimed one iteration ago
          publish(it + 2u, nxt);                                            // entries up to it + 2 are now visible
          tq1 = tq2; tq2 = nxt;
          if (nxt != 0xffffffffu) claim_issue();                             // for it + 3
          return cur != 0xffffffffu ? static_cast<int64_t>(cur) : n_tiles;
        }
        const int64_t t = tile_get(it, peeked);
        peeked = tile_peek(it + 1u);
        return t;
      };
      if (F32) {
        // fp32 boxes of 128 rows x 32 k (16 KB) into this CTA's own ring; the converter warps free a stage as soon as
        // its contents sit in their registers
        const uint32_t sF_a = smem_u32(sF), ffull_a = smem_u32(ffull_bar), fempty_a = smem_u32(fempty_bar);
        uint32_t stage = 0, phase = 0;
        for (int64_t t = first_tile(); t < n_tiles; t = next_tile(++tile_it)) {
          keep_in_step(tile_it);
          const int row0 = static_cast<int>(t * kTileRows + rank * 128);
#pragma unroll 1
          for (int kc = 0; kc < 16; ++kc) {
            mbar_wait(fempty_a + stage * 8u, phase ^ 1u);
            if (elect_one()) {
              mbar_arrive_expect_tx(ffull_a + stage * 8u, static_cast<uint32_t>(kStageBytes));
              tma_load_2d<1>(sF_a + stage * kStageBytes, &tm_bank, ffull_a + stage * 8u, kc * 32, row0, p.bank_hint);
            }
            if (++stage == static_cast<uint32_t>(p.n_fstages)) { stage = 0; phase ^= 1u; }
          }
        }
      } else {
        const uint32_t sA_a = smem_u32(sA), full_a = smem_u32(full_bar), empty_a = smem_u32(empty_bar);
        const uint32_t full_lead = (kCtas == 2) ? mapa_rank0(full_a) : full_a;
        uint32_t stage = 0, phase = 0;
        for (int64_t t = first_tile(); t < n_tiles; t = next_tile(++tile_it)) {
          keep_in_step(tile_it);
          const int row0 = static_cast<int>(t * kTileRows + rank * 128);
#pragma unroll 1
          for (int kc = 0; kc < 8; ++kc) {
            mbar_wait(empty_a + stage * 8u, phase ^ 1u);
            if (elect_one()) {
              if (rank == 0) mbar_arrive_expect_tx(full_a + stage * 8u, static_cast<uint32_t>(kStageBytes) * kCtas);
              tma_load_2d<kCtas>(sA_a + stage * kStageBytes, &tm_bank, full_lead + stage * 8u, kc * 64, row0, p.bank_hint);
            }
            if (++stage == static_cast<uint32_t>(p.n_stages)) { stage = 0; phase ^= 1u; }
          }
        }
      }
      if (lockstep && lane == 0) st_relaxed_gpu_u32(p.progress + pair_id, 0xffffffffu);   // done: nobody waits for this pair any more
    }
  } else if (warp == 1) {
    // ================================================================== MMA issuer (leader CTA)
    // Converged warp, one elected lane issues: a loop run by a single divergent lane makes the compiler wrap every
    // tcgen05 instruction in an elect/broadcast loop (~150 instructions per stage), and the issue thread, not the
    // tensor pipe, then bounds the tile period.
    if (rank == 0) {
      const uint32_t idesc = make_idesc_bf16(128 * kCtas, NB);
      mbar_wait(smem_u32(q_bar), 0);
      tc_fence_after();
      if (trace && lane == 0) trace[2] = globaltimer_ns();                         // query block resident
      const uint64_t a_base = make_smem_desc(smem_u32(sA)), b_base = make_smem_desc(smem_u32(sB));
      const uint32_t a_step = static_cast<uint32_t>(kStageBytes) >> 4, b_step = b_chunk_bytes >> 4;   // start-address field units
      const uint32_t full_a = smem_u32(full_bar), empty_a = smem_u32(empty_bar);
      const uint32_t tfull_a = smem_u32(tfull_bar), tempty_a = smem_u32(tempty_bar);
      uint32_t stage = 0, phase = 0, it = 0;
      unsigned long long peeked = tile_peek(1);
      for (int64_t t = t_first; t < n_tiles; ++it, t = tile_get(it, peeked), peeked = tile_peek(it + 1u)) {
        const uint32_t buf = it & 1u, bphase = (it >> 1) & 1u;
        mbar_wait(tempty_a + buf * 8u, bphase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * 256u;
#pragma unroll 1
        for (int kc = 0; kc < 8; ++kc) {
          mbar_wait(full_a + stage * 8u, phase);      // TMA transaction (bf16) / converter warps of both CTAs (F32)
          tc_fence_after();
          if (elect_one()) {
            const uint64_t a0 = a_base + stage * a_step;
            const uint64_t b0 = b_base + static_cast<uint32_t>(kc) * b_step;
#pragma unroll
            for (int k = 0; k < 4; ++k)   // 64-wide K chunk = 4 x UMMA_K(16); +32 B inside the swizzle atom
              umma_bf16<kCtas>(d_tmem, a0 + 2u * k, b0 + 2u * k, idesc, (kc | k) != 0 ? 1u : 0u);
            umma_commit<kCtas>(empty_a + stage * 8u);
          }
          if (++stage == static_cast<uint32_t>(p.n_stages)) { stage = 0; phase ^= 1u; }
        }
        if (elect_one()) umma_commit<kCtas>(tfull_a + buf * 8u);
        if (trace && it == 0 && lane == 0) trace[3] = globaltimer_ns();            // first tile's MMAs issued
      }
      if (trace && lane == 0) trace[4] = globaltimer_ns();                         // last tile's MMAs issued
    }
  } else if (warp == 3) {
    // ================================================================== threshold refresher
    // Recomputes the thresholds of this Q block's classes from their histograms, round-robin over the
    // CTAs serving the block, until this CTA's epilogue is done.  Keeps every returning atomic and
    // histogram read off the epilogue warps.
    if (!DENSE) {
      const int c_lo = p.blk_class[qb], c_hi = p.blk_class[qb + 1];
      const int n_ctas_qb = qb_peers * kCtas, my_idx = qb_rank * kCtas + static_cast<int>(rank);
      volatile uint32_t* done = s_done;
      const uint64_t t_start = globaltimer_ns();
      for (;;) {
        for (int c = c_lo + my_idx; c < c_hi; c += n_ctas_qb) refresh_tau(p.s.st, c);
        if (*done >= static_cast<uint32_t>(kEpiWarps)) break;
        if (globaltimer_ns() - t_start > 20000000000ull) __trap();
        __nanosleep(2000);
      }
    }
  } else if (F32 && warp >= 8) {
    // ================================================================== fp32 -> bf16 converter (F32 only)
    // Thread = one of the CTA's 128 tile rows.  A staged box holds 32 consecutive k of every row (128 B per row, the
    // eight 16-byte chunks XOR-swizzled with row & 7); two boxes fill one 64-k operand stage in the same swizzle.
    const int trow = (warp - 8) * 32 + lane;
    const uint32_t sw = static_cast<uint32_t>(trow & 7);
    const uint32_t src_row = smem_u32(sF) + static_cast<uint32_t>(trow) * 128u;
    const uint32_t dst_row = smem_u32(sA) + static_cast<uint32_t>(trow) * 128u;
    const uint32_t ffull_a = smem_u32(ffull_bar), fempty_a = smem_u32(fempty_bar);
    const uint32_t full_a = smem_u32(full_bar), empty_a = smem_u32(empty_bar);
    const uint32_t full_lead = (kCtas == 2) ? mapa_rank0(full_a) : full_a;
    uint32_t fstage = 0, fphase = 0, stage = 0, phase = 0, cit = 0;
    unsigned long long peeked = tile_peek(1);
    for (int64_t t = t_first; t < n_tiles; ++cit, t = tile_get(cit, peeked), peeked = tile_peek(cit + 1u)) {
#pragma unroll 1
      for (int kc = 0; kc < 16; ++kc) {
        mbar_wait(ffull_a + fstage * 8u, fphase);
        uint32_t w[16];
#pragma unroll
        for (uint32_t ch = 0; ch < 8; ++ch) {
          const float4 v = lds_f32x4(src_row + fstage * kStageBytes + ((ch ^ sw) << 4));
          w[2 * ch] = pack_bf16x2(v.x, v.y);
          w[2 * ch + 1] = pack_bf16x2(v.z, v.w);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive_local(fempty_a + fstage * 8u);      // the box is in registers: hand the stage back
        if (++fstage == static_cast<uint32_t>(p.n_fstages)) { fstage = 0; fphase ^= 1u; }
        const uint32_t half = static_cast<uint32_t>(kc & 1);
        if (half == 0) mbar_wait(empty_a + stage * 8u, phase ^ 1u);   // MMAs that read this operand stage have retired
#pragma unroll
        for (uint32_t ch = 0; ch < 4; ++ch)
          sts_u32x4(dst_row + stage * kStageBytes + (((half * 4u + ch) ^ sw) << 4), w[4 * ch], w[4 * ch + 1], w[4 * ch + 2], w[4 * ch + 3]);
        fence_proxy_async();          // generic-proxy writes -> visible to the tensor core's async-proxy reads
        __syncwarp();
        // Plain arrive on the leader's barrier (release at CTA scope): the tile lives in THIS CTA's shared memory and is
        // read by this SM's half of the tensor-core pair, so CTA-scope visibility plus the proxy fence is what the MMA
        // needs.  (A .release.cluster arrive compiles to MEMBAR.ALL.GPU + ERRBAR + CCTL.IVALL per stage: ncu showed the
        // converter warps spending most of their time in those, 2.2 TB/s instead of the HBM rate.)
        if (lane == 0) mbar_arrive_cluster(full_lead + stage * 8u);
        if (half == 1 && ++stage == static_cast<uint32_t>(p.n_stages)) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp >= 4) {
    // ================================================================== epilogue
    const int ew = warp - 4;          // 0..kEpiWarps-1
    const int quad = ew & 3;          // TMEM lanes 32*quad .. 32*quad+31 (warp id % 4)
    const int half = ew >> 2;         // which interleaved set of 32-column chunks
    const int etid = threadIdx.x - 128;
    EpiCtx cx;
    cx.cls_col = s_cls;
    cx.cnt_col = s_cnt;
    const uint32_t list_id = blockIdx.x * static_cast<uint32_t>(kTcEpiWarps) + static_cast<uint32_t>(ew);   // private to this warp
    const SlowCtx sc = make_slow_ctx(p.s, DENSE ? 0u : list_id);
    cx.list_pos = DENSE ? 0u : p.s.st.list_count[list_id];
    cx.tau_col = s_tau;
    // the epilogue threads own the columns of the threshold table between them: the values for the NEXT tile are
    // fetched while the current tile is processed
    constexpr int kColsPer = 256 / (kEpiWarps * 32);
    bool col_live[kColsPer];
    const uint32_t* tau_src[kColsPer];
    float col_cnt[kColsPer];
    uint32_t tnext[kColsPer];
#pragma unroll
    for (int i = 0; i < kColsPer; ++i) {
      const int col = etid + i * kEpiWarps * 32;
      col_live[i] = !DENSE && col < NB && s_cls[col] >= 0 && s_cnt[col] > 0.0f;
      tau_src[i] = &p.s.st.tau_enc[col_live[i] ? s_cls[col] : 0];
      col_cnt[i] = col_live[i] ? s_cnt[col] : 0.0f;
      tnext[i] = col_live[i] ? ld_cg_u32(tau_src[i]) : 0u;
    }
    uint32_t it = 0;
    unsigned long long peeked = 0;
    for (int64_t t = t_first; t < n_tiles; ++it, t = tile_get(it, peeked)) {
      const uint32_t buf = it & 1u, bphase = (it >> 1) & 1u;
      peeked = tile_peek(it + 1u);              // consumed after this tile: the load is in flight while the tile is processed
#pragma unroll
      for (int i = 0; i < kColsPer; ++i) {
        if (col_live[i]) {
          s_tau[etid + i * kEpiWarps * 32] = fast_tau<RED>(f32_dec(tnext[i]), col_cnt[i]);
          tnext[i] = ld_cg_u32(tau_src[i]);
        }
      }
      const int64_t row = t * kTileRows + rank * 128 + quad * 32 + lane;
      cx.row = static_cast<uint32_t>(row);
      cx.row_valid = row < p.s.n_rows && !row_excluded(p.s.exclude, static_cast<uint32_t>(row));
      cx.my_cls = -1;
      if (PART && cx.row_valid) cx.my_cls = p.s.row_class[row];
      cx.acc = red_init<RED>();
      cx.acc2 = 0.0f;
      mbar_wait(smem_u32(&tfull_bar[buf]), bphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + buf * 256u;
      const uint32_t tempty_a = smem_u32(&tempty_bar[buf]);
      const uint32_t tempty_lead = (kCtas == 2) ? mapa_rank0(tempty_a) : tempty_a;
      // 32 columns per TMEM load, double-buffered: the load of chunk i+1 is in flight while chunk i
      // is processed.  A buffer is 256 columns wide, so a trailing partial chunk simply reads columns
      // whose threshold is +inf.  The accumulator is handed back to the MMA warp as soon as its last
      // column sits in registers.
      auto release_tmem = [&]() {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(tempty_lead);
      };
      uint32_t ra[32], rb[32];
      float v[32];
      // One-query-per-class: the two warps of a quadrant take interleaved 32-column chunks.
      // Grouped reduces walk columns in order, so the block is cut at a class boundary (`split`,
      // a multiple of 32 chosen by the host) and each warp takes one contiguous part.
      constexpr bool kInterleave = (RED == RED_NONE) && kEpiWarps == 8;
      const int split = (kInterleave || kEpiWarps == 4) ? NB : p.blk_split[qb];
      const int first = kInterleave ? 32 * half : (half == 0 ? 0 : split);
      const int last = kInterleave ? NB : (half == 0 ? split : NB);
      constexpr int kStep = kInterleave ? 64 : 32;
      // Grouped reduces on 4 epilogue warps (fp32 banks): one warp walks the whole block, and the zero-query padding
      // columns the host inserts in front of `blk_split` (so that no class straddles it) would leak a 0 into the next
      // class's max / min: restart the running reduce where the second half begins.
      const int restart = (kEpiWarps == 4 && RED != RED_NONE) ? p.blk_split[qb] : -1;
      if (first < last) {
        tmem_ld32_issue(taddr + first, ra);
        tmem_wait(ra);
        for (int c0 = first; c0 < last; c0 += 2 * kStep) {
          const bool has_b = c0 + kStep < last, has_next = c0 + 2 * kStep < last;
          if (has_b) tmem_ld32_issue(taddr + c0 + kStep, rb); else release_tmem();
          to_f32x32(ra, v);
          if (c0 == restart) cx.acc = red_init<RED>();
          if (NB - c0 == 16) {   // trailing half chunk: columns the MMA never wrote must not look like scores
#pragma unroll
            for (int j = 16; j < 32; ++j) v[j] = 0.0f;
          }
          process_chunk<32, RED, PART, false, DENSE, false>(p.s, sc, cx, v, v, c0, s_end[c0 >> 5]);
          if (has_b) {
            tmem_wait(rb);
            if (has_next) tmem_ld32_issue(taddr + c0 + 2 * kStep, ra); else release_tmem();
            to_f32x32(rb, v);
            if (c0 + kStep == restart) cx.acc = red_init<RED>();
            if (NB - (c0 + kStep) == 16) {
#pragma unroll
              for (int j = 16; j < 32; ++j) v[j] = 0.0f;
            }
            process_chunk<32, RED, PART, false, DENSE, false>(p.s, sc, cx, v, v, c0 + kStep, s_end[(c0 + kStep) >> 5]);
            if (has_next) tmem_wait(ra);
          }
        }
      } else {
        release_tmem();   // nothing to read for this warp in this tile
      }
      if (trace && it == 0 && ew == 0 && lane == 0) trace[5] = globaltimer_ns();   // first tile through the epilogue
    }
    if (trace && ew == 0 && lane == 0) trace[6] = globaltimer_ns();                // last tile through the epilogue
    if (lane == 0) {
      if (!DENSE) p.s.st.list_count[list_id] = cx.list_pos;
      atomicAdd(s_done, 1u);
    }
  }

  tc_fence_before();
  if (kCtas == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 2) tmem_dealloc<kCtas>(tmem_base, 512);
  if (trace && threadIdx.x == 0) trace[7] = globaltimer_ns();                      // teardown
  if (probing) {
    const uint64_t ns = globaltimer_ns() - probe_t0;
    const uint64_t ticks = static_cast<uint64_t>(clock64() - probe_c0);
    if (ns > 200000ull) *p.clock_probe = ticks * 1000ull / ns;                     // MHz; launches under 0.2 ms say nothing
  }
}

template <int kCtas, int RED, bool PART, bool DENSE, bool F32>
cudaError_t launch_one(const CUtensorMap& tm_bank, const CUtensorMap& tm_q, const TcArgs& p, int grid, size_t smem, cudaStream_t stream) {
  auto kern = scan_tc_kernel<kCtas, RED, PART, DENSE, F32>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCtas;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, tm_bank, tm_q, p);
}

template <int kCtas, int RED, bool F32>
cudaError_t launch_red(const CUtensorMap& a, const CUtensorMap& b, const TcArgs& p, bool part, bool dense, int grid, size_t smem, cudaStream_t s) {
  if (dense) return launch_one<kCtas, RED, false, true, F32>(a, b, p, grid, smem, s);
  if (part) return launch_one<kCtas, RED, true, false, F32>(a, b, p, grid, smem, s);
  return launch_one<kCtas, RED, false, false, F32>(a, b, p, grid, smem, s);
}
template <int kCtas, bool F32>
cudaError_t launch_ctas(const CUtensorMap& a, const CUtensorMap& b, const TcArgs& p, int red, bool part, bool dense, int grid, size_t smem, cudaStream_t s) {
  switch (red) {
    case RED_NONE: return launch_red<kCtas, RED_NONE, F32>(a, b, p, part, dense, grid, smem, s);
    case RED_MEAN: return launch_red<kCtas, RED_MEAN, F32>(a, b, p, part, dense, grid, smem, s);
    case RED_MAX: return launch_red<kCtas, RED_MAX, F32>(a, b, p, part, dense, grid, smem, s);
    default: return launch_red<kCtas, RED_MIN, F32>(a, b, p, part, dense, grid, smem, s);
  }
}

}  // namespace

size_t tc_smem_bytes(int n_blk, int ctas, int n_stages, int n_fstages) {
  const size_t b = static_cast<size_t>(8) * (n_blk / ctas) * 128;
  return 1024 /* alignment slack */ + b + static_cast<size_t>(n_stages + n_fstages) * kStageBytes + kTailBytes;
}

int tc_pick_stages(int n_blk, int ctas, size_t smem_limit) {
  for (int s = 8; s >= 2; --s)
    if (tc_smem_bytes(n_blk, ctas, s, 0) <= smem_limit) return s;
  return 0;
}

// fp32 banks: `op` operand stages (64 k of bf16 each) + the returned number of staged fp32 boxes (32 k each)
int tc_pick_fstages(int n_blk, int ctas, size_t smem_limit, int* op_stages, int prefer_op) {
  for (int op = prefer_op; op >= 2; --op)
    for (int f = 8; f >= (op >= 3 ? 4 : 2); --f)
      if (tc_smem_bytes(n_blk, ctas, op, f) <= smem_limit) { *op_stages = op; return f; }
  *op_stages = 0;
  return 0;
}

cudaError_t launch_scan_tc(const void* tm_bank, const void* tm_q, const TcArgs& p, int ctas, int reduce, bool partitioned,
                           bool dense, bool f32, int grid, cudaStream_t stream) {
  const size_t smem = tc_smem_bytes(p.n_blk, ctas, p.n_stages, f32 ? p.n_fstages : 0);
  const CUtensorMap& a = *static_cast<const CUtensorMap*>(tm_bank);
  const CUtensorMap& b = *static_cast<const CUtensorMap*>(tm_q);
  if (f32) {
    if (ctas == 2) return launch_ctas<2, true>(a, b, p, reduce, partitioned, dense, grid, smem, stream);
    return launch_ctas<1, true>(a, b, p, reduce, partitioned, dense, grid, smem, stream);
  }
  if (ctas == 2) return launch_ctas<2, false>(a, b, p, reduce, partitioned, dense, grid, smem, stream);
  return launch_ctas<1, false>(a, b, p, reduce, partitioned, dense, grid, smem, stream);
}

}  // namespace swat
