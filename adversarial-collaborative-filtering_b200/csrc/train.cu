// APR / BPR-MF training step for sm_100a.
//
// Replaces, per batch (utils.py:113-119):
//     sess.run([update_P, update_Q])   APR.py:180-191   Delta_r = eps * G_r / ||G_r||, G = plain BPR gradient,
//                                                       duplicates of a row inside the batch summed
//     sess.run(optimizer)              APR.py:143-165,193-195   Adagrad on L(Theta) + reg_adv L(Theta+Delta) + reg
//
// Design (DESIGN.md "Training step"):
//   prepare   hash de-duplication of the rows a batch touches, for all steps of a chunk at once.
//             Users: every distinct user of a batch becomes a SEGMENT (its triples made contiguous by a counting
//             sort) owned by one lane group, so user-side sums (G_P, g_P) live in registers.
//             Items: an item that occurs ONCE in the batch is a singleton -- its row sums are local to its triple.
//             Only SHARED items (>= 2 occurrences) get a slot in the L2-resident workspace G_Q / H_Q [<= B, d].
//             Segments are PARTITIONED: general ones (touching a shared item, or several triples of one user) first,
//             fast ones at the end; in between, the segments taken by PAIR work units (below).
//   FAST segment (one triple, no shared item; ~88% at B=65536 on 10M x 2M): the whole APR step in registers -- gather
//             p, q, n and the three Adagrad rows with 128-bit loads, x, c = -sigma(-x), Delta from the local gradients,
//             adversarial forward (closed form), total gradient, Adagrad, six row stores.  HBM traffic is exactly the
//             algorithmic 4 row-transfers per touched row; nothing else is written.
//   PAIR work unit (~10% of the segments): two single-triple segments whose only coupling to the rest of the batch is
//             one item that occurs in both and nowhere else -- both triples in registers, the shared item's gradient
//             sums are two-term sums (pair_unit).  No workspace, no RED, no stage split.
//   GENERAL segments (the remaining ~2%), three stages separated by grid-wide ordering:
//     stage 0  plain forward: G_P kept, c kept per triple, G_Q[slot] += +-c p by 16-byte vector RED at L2.
//     stage 1  Delta_P from G_P, Delta_Q from G_Q[slot] (shared) or c p (singleton), adversarial forward; user row and
//              singleton item rows get Adagrad directly; H_Q[slot] += g (RED).
//     stage 2  per shared slot: Adagrad on the item row from H_Q[slot]; slots re-zeroed.
//   BPR (adver=0) has no stage 0.  Delta and per-triple gradients never exist in HBM as tables.
//   mode 0: per step one fast kernel (second stream), one pair kernel (third stream) and one kernel per general stage
//           (caller's stream), forked from / joined to the caller's stream; index preparation runs one L2-sized
//           sub-chunk ahead on a fourth, low-priority stream.
//   mode 1: all steps in ONE persistent cooperative kernel with grid barriers between stages.
//   mode 2: the same persistent kernel launched as ONE thread-block cluster (<= 16 CTAs, cluster barriers): batches of
//           a few hundred triples whose steps take a few microseconds.
//   Row-sharded tables (nranks > 1): every row access goes through shard_row(); a segment / pair is processed by the
//   rank that owns its (first) user row; apr_train_steps_sharded adds the cross-rank barriers.
#include <cooperative_groups.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "context.cuh"

namespace cg = cooperative_groups;

namespace apr {

static inline int pow2ceil(int x) {
  int p = 1;
  while (p < x) p <<= 1;
  return p;
}

// ------------------------------------------------------------------------------------------------
// workspace layout
// ------------------------------------------------------------------------------------------------
// Index preparation runs in sub-chunks of `Sc` steps so that its hash tables stay L2-resident (<= ~48 MB).
struct TrainLayout {
  int S, B, d, Tu, Ti, Sc;
  int64_t off_hdr, off_GQ, off_HQ, off_GP, off_cbuf, off_QN, off_ucnt, off_icnt, off_iall, off_nslow, off_nfast, off_tcursor,
      off_npair, off_seg_slow, off_occ_cnt, off_occ_seg, off_pairs, off_seg_user, off_seg_off, off_seg_cnt, off_ucursor, off_entry, off_rec, off_seg_hdr, off_iu_item, off_tkey_u, off_tval_u,
      off_tkey_i, off_tval_i, total;
};

static TrainLayout make_layout(int S, int B, int d) {
  TrainLayout L;
  L.S = S; L.B = B; L.d = d;
  L.Tu = std::max(32, 2 * pow2ceil(B));
  L.Ti = std::max(32, 4 * pow2ceil(B));
  const int64_t table_bytes_per_step = int64_t(L.Tu + L.Ti) * 8;
  L.Sc = int(std::max<int64_t>(1, std::min<int64_t>(S, (int64_t(48) << 20) / table_bytes_per_step)));
  int64_t o = 0;
  auto take = [&](int64_t bytes) { int64_t r = o; o += (bytes + 255) & ~int64_t(255); return r; };
  L.off_hdr = take(256);                       // word 0: id-out-of-range flag
  L.off_GQ = take(int64_t(B) * d * 4);         // zero between steps
  L.off_HQ = take(int64_t(B) * d * 4);         // zero between steps
  L.off_GP = take(int64_t(B) * d * 4);
  L.off_cbuf = take(int64_t(B) * 4);
  L.off_QN = take(int64_t(B) * 2 * d * 4);     // row-sharded tables: the two item rows of a stage-0 triple, kept for stage 1
  L.off_ucnt = take(int64_t(S) * 4);           // ucnt .. seg_slow are cleared together by prepare
  L.off_icnt = take(int64_t(S) * 4);
  L.off_iall = take(int64_t(S) * 4);
  L.off_nslow = take(int64_t(S) * 4);
  L.off_nfast = take(int64_t(S) * 4);
  L.off_tcursor = take(int64_t(S) * 4);
  L.off_npair = take(int64_t(S) * 4);
  L.off_seg_slow = take(int64_t(S) * B * 4);   // shared-item occurrences of the segment; -1 once it is part of a pair
  L.off_occ_cnt = take(int64_t(S) * B * 4);    // occurrences of a shared slot (cleared with the counters above)
  L.off_seg_user = take(int64_t(S) * B * 4);
  L.off_seg_off = take(int64_t(S) * B * 4);
  L.off_seg_cnt = take(int64_t(S) * B * 4);
  L.off_ucursor = take(int64_t(S) * B * 4);
  L.off_entry = take(int64_t(S) * B * 12);     // hash entries of (u, i, j) per triple
  L.off_rec = take(int64_t(S) * B * 16);       // {i, j, slot_i, slot_j} in segment order
  L.off_seg_hdr = take(int64_t(S) * B * 32);   // per segment {user, b0, count, slow} + its first triple's record,
                                               // partitioned: slow segments first, then fast ones
  L.off_iu_item = take(int64_t(S) * B * 4);    // shared slot -> item id (sign bit set: slot handled by the pair path)
  L.off_occ_seg = take(int64_t(S) * B * 8);    // first two segments that touch a shared slot
  L.off_pairs = take(int64_t(S) * (B / 2 + 1) * 48);  // pair work units: {user_a, user_b, slot, 0}, rec_a, rec_b
  L.off_tkey_u = take(int64_t(L.Sc) * L.Tu * 4);
  L.off_tval_u = take(int64_t(L.Sc) * L.Tu * 4);
  L.off_tkey_i = take(int64_t(L.Sc) * L.Ti * 4);
  L.off_tval_i = take(int64_t(L.Sc) * L.Ti * 4);
  L.total = o;
  return L;
}

// The layout of a workspace is a function of its SIZE (and of batch, d), not of the step count of the call that uses it:
// the capacity is the largest step count whose arrays fit, and a call of n <= capacity steps uses the first n step slots.
// So every array keeps its address from call to call -- which is what lets an executable CUDA graph of the step be
// replayed for calls of any length (the addresses are kernel parameters baked into its nodes).
static bool layout_for_ws(int64_t ws_bytes, int B, int d, int steps_needed, TrainLayout& L) {
  thread_local struct { int64_t bytes = -1; int B = 0, d = 0; TrainLayout L; } memo;
  if (!(memo.bytes == ws_bytes && memo.B == B && memo.d == d)) {
    if (make_layout(1, B, d).total > ws_bytes) return false;
    int64_t lo = 1, hi = 2;                                  // layout(lo) fits; total is monotone in the step count
    while (hi <= (int64_t(1) << 22) && make_layout(int(hi), B, d).total <= ws_bytes) { lo = hi; hi *= 2; }
    while (lo + 1 < hi) {
      const int64_t mid = (lo + hi) / 2;
      if (make_layout(int(mid), B, d).total <= ws_bytes) lo = mid; else hi = mid;
    }
    memo.bytes = ws_bytes; memo.B = B; memo.d = d;
    memo.L = make_layout(int(lo), B, d);
  }
  L = memo.L;
  return L.S >= steps_needed;
}

template <typename T>
static inline T* at(void* ws, int64_t off) { return reinterpret_cast<T*>(static_cast<char*>(ws) + off); }

static int env_int(const char* name, int dflt);
// pair work units (prep_pair_kernel / pair_unit): rows of up to 128 floats (one float4 per lane), single GPU
static bool pairs_enabled(int d) {
  static const int on = env_int("APR_PAIRS", 1);
  return on != 0 && d <= 128;
}

// ------------------------------------------------------------------------------------------------
// prepare kernels (all steps of a sub-chunk in parallel; one hash table per step).  `s0` = first step of the
// sub-chunk: per-step arrays are indexed by the absolute step, tables by the step within the sub-chunk.
// ------------------------------------------------------------------------------------------------
// continue an insertion whose first probe (slot h) returned `prev`
__device__ __forceinline__ uint32_t hash_resolve(int32_t* table, int T, int32_t key, uint32_t h, int32_t prev, bool& dup) {
  while (true) {
    if (prev == -1) { dup = false; return h; }
    if (prev == key) { dup = true; return h; }
    h = (h + 1) & uint32_t(T - 1);
    prev = atomicCAS(&table[h], -1, key);
  }
}

// one CAS per id; only duplicates pay a second atomic (users: extra count) or a flag store (items: "shared")
__global__ void prep_insert_kernel(const int32_t* __restrict__ u, const int32_t* __restrict__ i,
                                   const int32_t* __restrict__ j, int s0, int ns, int B, int64_t rows_p, int64_t rows_q,
                                   int Tu, int Ti, int32_t* tkey_u, int32_t* tval_u, int32_t* tkey_i, int32_t* tval_i,
                                   int32_t* entry, int32_t* hdr) {
  const int64_t total = int64_t(ns) * B;
  for (int64_t tl = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; tl < total; tl += int64_t(gridDim.x) * blockDim.x) {
    const int sl = int(tl / B);
    const int64_t t = int64_t(s0) * B + tl;
    int32_t ku = u[t], ki = i[t], kj = j[t];
    if (ku < 0 || ku >= rows_p) { atomicOr(hdr, 1); ku = 0; }
    if (ki < 0 || ki >= rows_q) { atomicOr(hdr, 1); ki = 0; }
    if (kj < 0 || kj >= rows_q) { atomicOr(hdr, 1); kj = 0; }
    // the three first probes go out together (independent L2 atomics in flight), collisions are resolved afterwards;
    // i and j share a table: the two CAS of one thread to one address stay in program order, so i == j makes j a duplicate
    int32_t* tu = tkey_u + int64_t(sl) * Tu;
    int32_t* ti = tkey_i + int64_t(sl) * Ti;
    uint32_t hu = fmix32(uint32_t(ku)) & uint32_t(Tu - 1);
    uint32_t hi = fmix32(uint32_t(ki)) & uint32_t(Ti - 1);
    uint32_t hj = fmix32(uint32_t(kj)) & uint32_t(Ti - 1);
    const int32_t pu = atomicCAS(&tu[hu], -1, ku);
    const int32_t pi = atomicCAS(&ti[hi], -1, ki);
    const int32_t pj = atomicCAS(&ti[hj], -1, kj);
    bool dup;
    hu = hash_resolve(tu, Tu, ku, hu, pu, dup);
    if (dup) atomicAdd(&tval_u[int64_t(sl) * Tu + hu], 1);
    hi = hash_resolve(ti, Ti, ki, hi, pi, dup);
    if (dup) tval_i[int64_t(sl) * Ti + hi] = 1;
    hj = hash_resolve(ti, Ti, kj, hj, pj, dup);
    if (dup) tval_i[int64_t(sl) * Ti + hj] = 1;
    entry[3 * t] = int32_t(hu); entry[3 * t + 1] = int32_t(hi); entry[3 * t + 2] = int32_t(hj);
  }
}

// one thread per hash-table entry; warps never straddle tables (Tu, Ti are multiples of 32).
// users: every entry -> segment id + a range [b0, b0+count) of the step's triple slots (one atomic per warp on the
// step's cursor; segment order is irrelevant).  items: flagged entries -> shared slot, others -> -1 (singleton).
__global__ void prep_compact_kernel(int s0, int ns, int B, int Tu, int Ti, const int32_t* __restrict__ tkey_u,
                                    int32_t* tval_u, const int32_t* __restrict__ tkey_i, int32_t* tval_i, int32_t* ucnt,
                                    int32_t* icnt, int32_t* iall, int32_t* tcursor, int32_t* seg_user, int32_t* seg_off,
                                    int32_t* seg_cnt, int32_t* ucursor, int32_t* iu_item) {
  const int64_t nU = int64_t(ns) * Tu, nI = int64_t(ns) * Ti;
  const unsigned lane = threadIdx.x & 31;
  for (int64_t base = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) & ~int64_t(31); base < nU + nI;
       base += int64_t(gridDim.x) * blockDim.x) {
    const int64_t e = base + lane;
    const bool is_user = e < nU;
    int32_t key = -1, val = 0;
    int s = 0;
    int64_t ei = 0;
    if (is_user) { s = s0 + int(e / Tu); key = tkey_u[e]; if (key != -1) val = tval_u[e]; }
    else { ei = e - nU; s = s0 + int(ei / Ti); key = tkey_i[ei]; if (key != -1) val = tval_i[ei]; }
    const bool valid = key != -1;
    const bool slotted = valid && (is_user || val != 0);
    const unsigned mv = __ballot_sync(0xffffffffu, valid);
    const unsigned ms = __ballot_sync(0xffffffffu, slotted);
    if (mv == 0) continue;
    // inclusive warp scan of the user segment sizes
    const int cnt = (is_user && valid) ? 1 + val : 0;
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, incl, o);
      if (int(lane) >= o) incl += y;
    }
    const int warp_total = __shfl_sync(0xffffffffu, incl, 31);
    int basev = 0, baseo = 0;
    const int leader = __ffs(mv) - 1;
    if (int(lane) == leader) {
      if (is_user) { basev = atomicAdd(&ucnt[s], __popc(ms)); baseo = atomicAdd(&tcursor[s], warp_total); }
      else { atomicAdd(&iall[s], __popc(mv)); if (ms) basev = atomicAdd(&icnt[s], __popc(ms)); }
    }
    basev = __shfl_sync(0xffffffffu, basev, leader);
    baseo = __shfl_sync(0xffffffffu, baseo, leader);
    if (valid) {
      const int slot = slotted ? basev + __popc(ms & ((1u << lane) - 1u)) : -1;
      if (is_user) {
        const int64_t k = int64_t(s) * B + slot;
        seg_user[k] = key;
        seg_off[k] = baseo + incl - cnt;
        seg_cnt[k] = cnt;
        ucursor[k] = 0;
        tval_u[e] = slot;
      } else {
        if (slot >= 0) iu_item[int64_t(s) * B + slot] = key;
        tval_i[ei] = slot;
      }
    }
  }
}

// Block-aggregated variant for tables that are multiples of 256 entries (batch >= 128): a 256-thread block covers 256
// consecutive entries of ONE table of ONE step, the eight warps' totals meet in shared memory, and the per-step
// counters take one atomic per block instead of one per warp -- those same-address atomics were what the kernel cost.
__global__ void __launch_bounds__(256)
prep_compact_block_kernel(int s0, int ns, int B, int Tu, int Ti, const int32_t* __restrict__ tkey_u, int32_t* tval_u,
                          const int32_t* __restrict__ tkey_i, int32_t* tval_i, int32_t* ucnt, int32_t* icnt, int32_t* iall,
                          int32_t* tcursor, int32_t* seg_user, int32_t* seg_off, int32_t* seg_cnt, int32_t* ucursor,
                          int32_t* iu_item) {
  __shared__ int sh_slot[8], sh_tot[8], sh_val[8], sh_base[2];
  const int64_t nU = int64_t(ns) * Tu, nI = int64_t(ns) * Ti;
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int64_t base = int64_t(blockIdx.x) * 256; base < nU + nI; base += int64_t(gridDim.x) * 256) {
    const int64_t e = base + threadIdx.x;
    const bool is_user = base < nU;                 // block-uniform
    int32_t key, val = 0;
    int s;
    int64_t ei = 0;
    if (is_user) { s = s0 + int(base / Tu); key = tkey_u[e]; if (key != -1) val = tval_u[e]; }
    else { ei = e - nU; s = s0 + int((base - nU) / Ti); key = tkey_i[ei]; if (key != -1) val = tval_i[ei]; }
    const bool valid = key != -1;
    const bool slotted = valid && (is_user || val != 0);
    const unsigned mv = __ballot_sync(0xffffffffu, valid);
    const unsigned ms = __ballot_sync(0xffffffffu, slotted);
    const int cnt = (is_user && valid) ? 1 + val : 0;
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, incl, o);
      if (int(lane) >= o) incl += y;
    }
    if (lane == 31) { sh_slot[warp] = __popc(ms); sh_tot[warp] = incl; sh_val[warp] = __popc(mv); }
    __syncthreads();
    if (threadIdx.x == 0) {
      int nslot = 0, ntot = 0, nval = 0;
#pragma unroll
      for (int w = 0; w < 8; ++w) {
        const int a = sh_slot[w], b = sh_tot[w];
        sh_slot[w] = nslot; sh_tot[w] = ntot;       // exclusive prefixes over the warps
        nslot += a; ntot += b; nval += sh_val[w];
      }
      int bslot = 0, boff = 0;
      if (nval) {
        if (is_user) { bslot = atomicAdd(&ucnt[s], nslot); boff = atomicAdd(&tcursor[s], ntot); }
        else { atomicAdd(&iall[s], nval); if (nslot) bslot = atomicAdd(&icnt[s], nslot); }
      }
      sh_base[0] = bslot; sh_base[1] = boff;
    }
    __syncthreads();
    if (valid) {
      const int slot = slotted ? sh_base[0] + sh_slot[warp] + __popc(ms & ((1u << lane) - 1u)) : -1;
      if (is_user) {
        const int64_t k = int64_t(s) * B + slot;
        seg_user[k] = key;
        seg_off[k] = sh_base[1] + sh_tot[warp] + incl - cnt;
        seg_cnt[k] = cnt;
        ucursor[k] = 0;
        tval_u[e] = slot;
      } else {
        if (slot >= 0) iu_item[int64_t(s) * B + slot] = key;
        tval_i[ei] = slot;
      }
    }
    __syncthreads();   // sh_* are rewritten by the next iteration
  }
}

__global__ void prep_scatter_kernel(const int32_t* __restrict__ i, const int32_t* __restrict__ j, int s0, int ns, int B,
                                    int Tu, int Ti, const int32_t* __restrict__ tval_u, const int32_t* __restrict__ tval_i,
                                    const int32_t* __restrict__ entry, const int32_t* __restrict__ seg_off,
                                    int32_t* ucursor, int4* rec, int32_t* seg_slow, int32_t* occ_cnt, int2* occ_seg) {
  const int64_t total = int64_t(ns) * B;
  for (int64_t tl = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; tl < total; tl += int64_t(gridDim.x) * blockDim.x) {
    const int sl = int(tl / B);
    const int s = s0 + sl;
    const int64_t t = int64_t(s0) * B + tl;
    const int useg = tval_u[int64_t(sl) * Tu + entry[3 * t]];
    const int si = tval_i[int64_t(sl) * Ti + entry[3 * t + 1]];
    const int sj = tval_i[int64_t(sl) * Ti + entry[3 * t + 2]];
    const int pos = seg_off[int64_t(s) * B + useg] + atomicAdd(&ucursor[int64_t(s) * B + useg], 1);
    rec[int64_t(s) * B + pos] = make_int4(i[t], j[t], si, sj);
    const int nsh = (si >= 0 ? 1 : 0) + (sj >= 0 ? 1 : 0);
    if (nsh) {
      atomicAdd(&seg_slow[int64_t(s) * B + useg], nsh);
      // the first two segments that touch a shared slot (pair detection); a triple with i == j registers twice
      if (occ_cnt && si >= 0) {
        const int k = atomicAdd(&occ_cnt[int64_t(s) * B + si], 1);
        if (k < 2) reinterpret_cast<int32_t*>(&occ_seg[int64_t(s) * B + si])[k] = useg;
      }
      if (occ_cnt && sj >= 0) {
        const int k = atomicAdd(&occ_cnt[int64_t(s) * B + sj], 1);
        if (k < 2) reinterpret_cast<int32_t*>(&occ_seg[int64_t(s) * B + sj])[k] = useg;
      }
    }
  }
}

// PAIR detection, one thread per shared slot.  In a uniform batch ~97 % of the shared items occur exactly twice, in two
// different single-triple segments that have no other shared item: such a component of two segments is one PAIR work
// unit, processed by one group entirely in registers (pair_unit) -- no workspace slot, no RED, no stage split.  Each
// eligible segment has exactly ONE shared occurrence, so it belongs to at most one slot: no two threads claim it.
__global__ void prep_pair_kernel(int s0, int ns, int B, const int32_t* __restrict__ icnt,
                                 const int32_t* __restrict__ occ_cnt, const int2* __restrict__ occ_seg,
                                 const int32_t* __restrict__ seg_user, const int32_t* __restrict__ seg_off,
                                 const int32_t* __restrict__ seg_cnt, int32_t* seg_slow, const int4* __restrict__ rec,
                                 int32_t* iu_item, int32_t* npair, int4* pairs) {
  const int64_t total = int64_t(ns) * B;
  for (int64_t tl = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; tl < total; tl += int64_t(gridDim.x) * blockDim.x) {
    const int s = s0 + int(tl / B);
    const int slot = int(tl - int64_t(s - s0) * B);
    if (slot >= icnt[s]) continue;
    const int64_t b = int64_t(s) * B;
    if (occ_cnt[b + slot] != 2) continue;
    const int2 ab = occ_seg[b + slot];
    if (ab.x == ab.y) continue;                                   // i == j inside one triple
    if (seg_cnt[b + ab.x] != 1 || seg_cnt[b + ab.y] != 1) continue;
    if (seg_slow[b + ab.x] != 1 || seg_slow[b + ab.y] != 1) continue;
    // one atomic per (warp, step) instead of one per pair: lanes of the same step allocate together
    const unsigned peers = __match_any_sync(__activemask(), s);
    const int leader = __ffs(peers) - 1;
    int k = 0;
    if (int(threadIdx.x & 31) == leader) k = atomicAdd(&npair[s], __popc(peers));
    k = __shfl_sync(peers, k, leader) + __popc(peers & ((1u << (threadIdx.x & 31)) - 1u));
    int4* out = pairs + (int64_t(s) * (B / 2 + 1) + k) * 3;
    out[0] = make_int4(seg_user[b + ab.x], seg_user[b + ab.y], slot, 0);
    out[1] = rec[b + seg_off[b + ab.x]];
    out[2] = rec[b + seg_off[b + ab.y]];
    seg_slow[b + ab.x] = -1;
    seg_slow[b + ab.y] = -1;
    iu_item[b + slot] |= int32_t(0x80000000u);                    // stage 2 skips this slot
  }
}

// one thread per segment: packed 32-byte header, written in PARTITIONED order -- GENERAL segments (touching a shared
// item, or holding several triples of one user) occupy [0, ngen), FAST ones (one triple, both items singletons) the
// END of the range, [nu - nfast, nu); the 2 npair segments taken by pair work units get no header (the gap between the
// two) -- so each kernel walks a dense range.  header.w != 0 iff the segment touches a shared item.
__global__ void prep_pack_kernel(int s0, int ns, int B, const int32_t* __restrict__ ucnt,
                                 const int32_t* __restrict__ seg_user, const int32_t* __restrict__ seg_off,
                                 const int32_t* __restrict__ seg_cnt, const int32_t* __restrict__ seg_slow,
                                 const int4* __restrict__ rec, int4* seg_hdr, int32_t* nslow, int32_t* nfast) {
  const int64_t total = int64_t(ns) * B;
  const unsigned lane = threadIdx.x & 31;
  for (int64_t base = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) & ~int64_t(31); base < total;
       base += int64_t(gridDim.x) * blockDim.x) {
    const int64_t tl = base + lane;
    bool valid = tl < total;
    int s = 0, seg = 0, nu = 0;
    if (valid) { s = s0 + int(tl / B); seg = int(tl - int64_t(s - s0) * B); nu = ucnt[s]; valid = seg < nu; }
    const int64_t t = int64_t(s) * B + seg;
    const int shared = valid ? seg_slow[t] : 0;
    if (shared < 0) valid = false;   // part of a pair: neither in the general nor in the fast range
    int b0 = 0, b1 = 0;
    if (valid) { b0 = seg_off[t]; b1 = b0 + seg_cnt[t]; }
    const int slow = (shared || (b1 - b0) != 1) ? 1 : 0;
    // warp-aggregated counters, one aggregation per distinct step present in the warp (at most 2 when B >= 32)
    int k = -1;
    unsigned todo = __ballot_sync(0xffffffffu, valid);
    while (todo) {
      const int leader = __ffs(todo) - 1;
      const int ls = __shfl_sync(0xffffffffu, s, leader);
      const bool mine = valid && s == ls;
      const unsigned m_slow = __ballot_sync(0xffffffffu, mine && slow);
      const unsigned m_fast = __ballot_sync(0xffffffffu, mine && !slow);
      int bs = 0, bf = 0;
      if (int(lane) == leader) {
        if (m_slow) bs = atomicAdd(&nslow[ls], __popc(m_slow));
        if (m_fast) bf = atomicAdd(&nfast[ls], __popc(m_fast));
      }
      bs = __shfl_sync(0xffffffffu, bs, leader);
      bf = __shfl_sync(0xffffffffu, bf, leader);
      if (mine) {
        const unsigned lt = (1u << lane) - 1u;
        k = slow ? bs + __popc(m_slow & lt) : nu - 1 - (bf + __popc(m_fast & lt));
      }
      todo &= ~__ballot_sync(0xffffffffu, mine);
    }
    if (valid) {
      int4* h = seg_hdr + (int64_t(s) * B + k) * 2;
      h[0] = make_int4(seg_user[t], b0, b1 - b0, shared);
      h[1] = rec[int64_t(s) * B + b0];
    }
  }
}

// Block-aggregated variant for batch % 256 == 0: a block covers 256 segments of ONE step, so the two per-step counters
// take one atomic per block.
__global__ void __launch_bounds__(256)
prep_pack_block_kernel(int s0, int ns, int B, const int32_t* __restrict__ ucnt, const int32_t* __restrict__ seg_user,
                       const int32_t* __restrict__ seg_off, const int32_t* __restrict__ seg_cnt,
                       const int32_t* __restrict__ seg_slow, const int4* __restrict__ rec, int4* seg_hdr, int32_t* nslow,
                       int32_t* nfast) {
  __shared__ int sh_s[8], sh_f[8], sh_base[2];
  const int64_t total = int64_t(ns) * B;
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int64_t base = int64_t(blockIdx.x) * 256; base < total; base += int64_t(gridDim.x) * 256) {
    const int s = s0 + int(base / B);                       // block-uniform
    const int seg = int(base - int64_t(s - s0) * B) + threadIdx.x;
    const int nu = ucnt[s];
    bool valid = seg < nu;
    const int64_t t = int64_t(s) * B + seg;
    const int shared = valid ? seg_slow[t] : 0;
    if (shared < 0) valid = false;                          // part of a pair
    int b0 = 0, b1 = 0;
    if (valid) { b0 = seg_off[t]; b1 = b0 + seg_cnt[t]; }
    const int slow = (shared || (b1 - b0) != 1) ? 1 : 0;
    const unsigned m_slow = __ballot_sync(0xffffffffu, valid && slow);
    const unsigned m_fast = __ballot_sync(0xffffffffu, valid && !slow);
    if (lane == 0) { sh_s[warp] = __popc(m_slow); sh_f[warp] = __popc(m_fast); }
    __syncthreads();
    if (threadIdx.x == 0) {
      int ts = 0, tf = 0;
#pragma unroll
      for (int w = 0; w < 8; ++w) {
        const int a = sh_s[w], b = sh_f[w];
        sh_s[w] = ts; sh_f[w] = tf;
        ts += a; tf += b;
      }
      sh_base[0] = ts ? atomicAdd(&nslow[s], ts) : 0;
      sh_base[1] = tf ? atomicAdd(&nfast[s], tf) : 0;
    }
    __syncthreads();
    if (valid) {
      const unsigned lt = (1u << lane) - 1u;
      const int k = slow ? sh_base[0] + sh_s[warp] + __popc(m_slow & lt)
                         : nu - 1 - (sh_base[1] + sh_f[warp] + __popc(m_fast & lt));
      int4* h = seg_hdr + (int64_t(s) * B + k) * 2;
      h[0] = make_int4(seg_user[t], b0, b1 - b0, shared);
      h[1] = rec[int64_t(s) * B + b0];
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// step kernels
// ------------------------------------------------------------------------------------------------
constexpr int kMaxRanks = 8;
// The per-launch part of a step's arguments, kept in DEVICE memory (workspace header, byte 64) so that an executable CUDA
// graph of the step can be replayed without touching its kernel parameters: `s` = absolute step the next graph launch
// starts at (advanced by the graph's last node), `stats` = the caller's per-step {loss, correct} array (nullable).
struct StepDyn { int s; int pad; float* stats; };
constexpr int64_t kDynOffset = 64;
// Tables are row-sharded: row r lives on rank r & (nranks-1) at local row r >> rshift (nranks a power of two; 1 on a
// single GPU).  Pb/Qb/... hold every rank's shard base (peer-mapped over NVLink); G_Q/H_Q slots are sharded the same way.
struct StepCtx {
  float* Pb[kMaxRanks]; float* Qb[kMaxRanks]; float* aPb[kMaxRanks]; float* aQb[kMaxRanks];
  float* GQb[kMaxRanks]; float* HQb[kMaxRanks];
  int nranks, rank, rshift;
  int d, B, S;
  float lr, kreg, reg_adv, eps;
  float plain_scale;   // factor on the plain data term of a NON-adversarial step: 1, or 1 + reg_adv for the reference's
                       // dns > 1 branch on an adversarial graph (utils.py:121-139: optimizer on opt_loss with Delta == 0)
  uint32_t noise_seed, noise_step0;   // --adv random: Philox key / first global step of this call (random_step_kernel)
  const int* abort;    // nullable: a cross-rank barrier timed out (sharded path) -> every later kernel of the call is a no-op
  int adver;
  const int32_t* ucnt; const int32_t* icnt; const int32_t* nslow;
  const int32_t* npair;   // nullable: pair work units per step (segments [nslow, nslow + 2 npair) are theirs)
  const int4* pairs;
  const int4* seg_hdr; const int4* rec; const int32_t* iu_item;
  float* GP; float* cbuf;
  float* QN;     // [B][2][d]: item rows fetched by stage 0 (row-sharded tables only: stage 1 re-reads them locally, not over NVLink)
  float* stats;  // nullable [S,2]
  const StepDyn* dyn;   // non-null inside a replayed CUDA graph: step = dyn->s + the kernel's step parameter, stats = dyn->stats
  int flags;     // tuning switches (APR_STEP_FLAGS)
  int s_begin, s_end;  // steps of this launch
  int only_stage;      // -1: whole step; 0,1,2: that general stage; 3: fast kernel (sharded driver, one launch per call)
  int cluster_sync;    // persistent kernel launched as ONE thread-block cluster: phase barriers are cluster barriers
};

__device__ __forceinline__ float* shard_row(float* const* base, int id, int d, const StepCtx& c) {
  return base[id & (c.nranks - 1)] + int64_t(id >> c.rshift) * d;
}
#define P_ROW(c, id) shard_row((c).Pb, (id), (c).d, (c))
#define Q_ROW(c, id) shard_row((c).Qb, (id), (c).d, (c))
#define AP_ROW(c, id) shard_row((c).aPb, (id), (c).d, (c))
#define AQ_ROW(c, id) shard_row((c).aQb, (id), (c).d, (c))
// APR_STEP_FLAGS bit 0 (TIMING EXPERIMENT ONLY, results are wrong): every access to the shared-item workspace goes to this
// rank's own shard, i.e. the vector REDs / reads of G_Q, H_Q stop crossing NVLink -- isolates what they cost
__device__ __forceinline__ float* slot_row(float* const* base, int slot, int d, const StepCtx& c) {
  const int owner = (c.flags & 1) ? c.rank : (slot & (c.nranks - 1));
  return base[owner] + int64_t(slot >> c.rshift) * d;
}
#define GQ_ROW(c, slot) slot_row((c).GQb, (slot), (c).d, (c))
#define HQ_ROW(c, slot) slot_row((c).HQb, (slot), (c).d, (c))

template <int G, int V>
struct Row {
  float4 v[V];
};

template <int G, int V>
__device__ __forceinline__ void row_load(Row<G, V>& r, const float* base, int lane, int d) {
#pragma unroll
  for (int k = 0; k < V; ++k) {
    const int e = (k * G + lane) * 4;
    r.v[k] = (e < d) ? ldcg4(base + e) : f4_zero();
  }
}
template <int G, int V>
__device__ __forceinline__ void row_store(const Row<G, V>& r, float* base, int lane, int d) {
#pragma unroll
  for (int k = 0; k < V; ++k) {
    const int e = (k * G + lane) * 4;
    if (e < d) stcg4(base + e, r.v[k]);
  }
}
template <int G, int V>
__device__ __forceinline__ void row_red(const Row<G, V>& r, float* base, int lane, int d) {
#pragma unroll
  for (int k = 0; k < V; ++k) {
    const int e = (k * G + lane) * 4;
    if (e < d) red_add4(base + e, r.v[k]);
  }
}
template <int G, int V>
__device__ __forceinline__ float row_dot(const Row<G, V>& a, const Row<G, V>& b, unsigned mask) {
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < V; ++k) s += f4_dot(a.v[k], b.v[k]);
  return group_sum<G>(s, mask);
}
template <int G, int V>
__device__ __forceinline__ void row_zero(Row<G, V>& r) {
#pragma unroll
  for (int k = 0; k < V; ++k) r.v[k] = f4_zero();
}

// single MUFU.RSQ (operands here are >= 1e-12, never denormal)
__device__ __forceinline__ float rsqrt_fast(float x) {
  float y;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ float softplus_neg(float r) {  // softplus(-r)
  return fmaxf(-r, 0.f) + log1pf(expf(-fabsf(r)));
}
// c = d softplus(-clip(x,-80,1e8)) / dx  (APR.py:148-150; the clip passes gradient on the closed interval)
__device__ __forceinline__ float bpr_coeff(float x, float& r_out) {
  const float r = fminf(fmaxf(x, -80.0f), 1e8f);
  r_out = r;
  const float m = (x >= -80.0f && x <= 1e8f) ? 1.0f : 0.0f;
  return __fdividef(-m, 1.0f + expf(r));
}

// Adagrad on one row with the accumulator row already in registers
template <int G, int V>
__device__ __forceinline__ void adagrad_apply(float* W, float* A, const Row<G, V>& w, Row<G, V> a, const Row<G, V>& g,
                                              int lane, int d, float lr) {
#pragma unroll
  for (int k = 0; k < V; ++k) {
    const int e = (k * G + lane) * 4;
    if (e < d) {
      const float4 gg = g.v[k];
      float4 aa = a.v[k];
      aa.x = fmaf(gg.x, gg.x, aa.x); aa.y = fmaf(gg.y, gg.y, aa.y); aa.z = fmaf(gg.z, gg.z, aa.z); aa.w = fmaf(gg.w, gg.w, aa.w);
      float4 nw = w.v[k];
      // w -= lr g / sqrt(a): MUFU.RSQ (2 ulp) instead of IEEE sqrt + divide -- 6x fewer issue slots, still ~1e-7
      nw.x = fmaf(-lr * gg.x, rsqrt_fast(aa.x), nw.x); nw.y = fmaf(-lr * gg.y, rsqrt_fast(aa.y), nw.y);
      nw.z = fmaf(-lr * gg.z, rsqrt_fast(aa.z), nw.z); nw.w = fmaf(-lr * gg.w, rsqrt_fast(aa.w), nw.w);
      stcg4(A + e, aa);
      stcg4(W + e, nw);
    }
  }
}
template <int G, int V>
__device__ __forceinline__ void adagrad_row(float* W, float* A, const Row<G, V>& w, const Row<G, V>& g, int lane, int d,
                                            float lr) {
  Row<G, V> a;
  row_load<G, V>(a, A, lane, d);
  adagrad_apply<G, V>(W, A, w, a, g, lane, d, lr);
}

// eps * rsqrt(max(||G||^2, 1e-12))  (tf.nn.l2_normalize epsilon, APR.py:190-191)
template <int G, int V>
__device__ __forceinline__ float delta_scale(const Row<G, V>& g, float eps, unsigned mask) {
  const float ss = row_dot<G, V>(g, g, mask);
  return eps * rsqrt_fast(fmaxf(ss, 1e-12f));
}

// block-level accumulation of {loss, correct} into stats[2*s .. 2*s+1]
__device__ __forceinline__ void stats_flush(float* stats, int s, float loss, float correct, bool group_leader) {
  __shared__ float sh[2];
  if (threadIdx.x == 0) { sh[0] = 0.f; sh[1] = 0.f; }
  __syncthreads();
  if (group_leader && (loss != 0.f || correct != 0.f)) { atomicAdd(&sh[0], loss); atomicAdd(&sh[1], correct); }
  __syncthreads();
  if (threadIdx.x == 0 && (sh[0] != 0.f || sh[1] != 0.f)) {
    atomicAdd(&stats[2 * s], sh[0]);
    atomicAdd(&stats[2 * s + 1], sh[1]);
  }
  __syncthreads();
}

// item-row gradient sink: shared slot -> RED into the workspace; singleton -> Adagrad right here
template <int G, int V>
__device__ __forceinline__ void item_sink(const StepCtx& c, int item, int slot, const Row<G, V>& w, const Row<G, V>& g,
                                          int lane, int d) {
  if (slot >= 0) row_red<G, V>(g, HQ_ROW(c, slot), lane, d);
  else adagrad_row<G, V>(Q_ROW(c, item), AQ_ROW(c, item), w, g, lane, d, c.lr);
}

// adversarial forward/backward of one triple given the plain coefficient cf and the row gradients gi, gj of its items
// (from the workspace when shared, +-cf*p when singleton).  Accumulates the user gradient, sinks the item gradients.
template <int G, int V>
__device__ __forceinline__ void adv_triple(const StepCtx& c, const int4 rc, float cf, const Row<G, V>& p,
                                           const Row<G, V>& pd, const Row<G, V>& q, const Row<G, V>& n, Row<G, V>& g,
                                           int lane, int d, unsigned mask) {
  Row<G, V> gi, gj;
  if (rc.z >= 0) row_load<G, V>(gi, GQ_ROW(c, rc.z), lane, d);
  else {
#pragma unroll
    for (int k = 0; k < V; ++k) gi.v[k] = f4_scale(p.v[k], cf);
  }
  if (rc.w >= 0) row_load<G, V>(gj, GQ_ROW(c, rc.w), lane, d);
  else {
#pragma unroll
    for (int k = 0; k < V; ++k) gj.v[k] = f4_scale(p.v[k], -cf);
  }
  const float sci = delta_scale<G, V>(gi, c.eps, mask);
  const float scj = delta_scale<G, V>(gj, c.eps, mask);
  Row<G, V> qd, nd;
#pragma unroll
  for (int k = 0; k < V; ++k) {
    qd.v[k] = f4_fma(sci, gi.v[k], q.v[k]);
    nd.v[k] = f4_fma(scj, gj.v[k], n.v[k]);
  }
  const float xa = row_dot<G, V>(pd, qd, mask) - row_dot<G, V>(pd, nd, mask);
  float ra;
  const float ca = c.reg_adv * bpr_coeff(xa, ra);
  Row<G, V> hi, hj;
#pragma unroll
  for (int k = 0; k < V; ++k) {
    g.v[k] = f4_fma(ca, f4_sub(qd.v[k], nd.v[k]), g.v[k]);
    g.v[k] = f4_fma(c.kreg, p.v[k], g.v[k]);
    const float4 t = f4_fma(ca, pd.v[k], f4_scale(p.v[k], cf));  // c p + reg_adv c' (p + dP)
    hi.v[k] = f4_fma(c.kreg, q.v[k], t);
    hj.v[k] = f4_fma(c.kreg, n.v[k], f4_scale(t, -1.f));
  }
  item_sink<G, V>(c, rc.x, rc.z, q, hi, lane, d);
  item_sink<G, V>(c, rc.y, rc.w, n, hj, lane, d);
}

struct StepStats { float loss, correct; bool on; };

template <int G>
__device__ __forceinline__ int4 shfl4(unsigned mask, const int4 v, int src) {
  int4 r;
  r.x = __shfl_sync(mask, v.x, src, G); r.y = __shfl_sync(mask, v.y, src, G);
  r.z = __shfl_sync(mask, v.z, src, G); r.w = __shfl_sync(mask, v.w, src, G);
  return r;
}

// ---------------------------------------------------------------------------------------------------------
// FAST segment: one triple, both items singletons.  Six row reads, six row writes, everything else in registers.
// The adversarial forward is evaluated in closed form from four row reductions (S_pq, S_pn, S_pp, S_dd with
// dq = q - n), which is the same arithmetic as APR.py:130-165 with the sums regrouped:
//     G_P = c dq, G_i = c p = -G_j             (plain gradients, APR.py:183-187)
//     a = eps c rsqrt(max(c^2 S_dd, 1e-12))    (Delta_P = a dq)       b = eps c rsqrt(max(c^2 S_pp, 1e-12))  (Delta_i = b p)
//     x' = <p + a dq, (q + b p) - (n - b p)> = x + 2 b S_pp + a S_dd + 2 a b x
//     g_P = (c + c') dq + (2 c' b + k) p       t = (c + c') p + c' a dq       g_i = t + k q       g_j = k n - t
// with c' = reg_adv * coeff(x').  FULL = row width equals d (no tail predicates).
// ---------------------------------------------------------------------------------------------------------
template <int G, int V, bool FULL>
__device__ __forceinline__ void fast_segment(const StepCtx& c, const int user, const int item_i, const int item_j, int lane,
                                             unsigned mask, StepStats& st) {
  const int d = c.d;
  float* Pu = P_ROW(c, user);
  float* Au = AP_ROW(c, user);
  float* Qi = Q_ROW(c, item_i);
  float* Ai = AQ_ROW(c, item_i);
  float* Qj = Q_ROW(c, item_j);
  float* Aj = AQ_ROW(c, item_j);
  float4 p[V], q[V], n[V], ap[V], ai[V], aj[V];
#pragma unroll
  for (int k = 0; k < V; ++k) {
    const int e = (k * G + lane) * 4;
    const bool ok = FULL || e < d;
    p[k] = ok ? ldcg4(Pu + e) : f4_zero();
    q[k] = ok ? ldcg4(Qi + e) : f4_zero();
    n[k] = ok ? ldcg4(Qj + e) : f4_zero();
    ap[k] = ok ? ldcg4(Au + e) : f4_zero();
    ai[k] = ok ? ldcg4(Ai + e) : f4_zero();
    aj[k] = ok ? ldcg4(Aj + e) : f4_zero();
  }
  float s_pq = 0.f, s_pn = 0.f, s_pp = 0.f, s_dd = 0.f;
#pragma unroll
  for (int k = 0; k < V; ++k) {
    s_pq += f4_dot(p[k], q[k]);
    s_pn += f4_dot(p[k], n[k]);
    s_pp += f4_dot(p[k], p[k]);
    const float4 dq = f4_sub(q[k], n[k]);
    s_dd += f4_dot(dq, dq);
  }
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) {  // four independent butterflies, interleaved
    s_pq += __shfl_xor_sync(mask, s_pq, o);
    s_pn += __shfl_xor_sync(mask, s_pn, o);
    s_pp += __shfl_xor_sync(mask, s_pp, o);
    s_dd += __shfl_xor_sync(mask, s_dd, o);
  }
  const float x = s_pq - s_pn;
  float r;
  const float cf = bpr_coeff(x, r);
  if (st.on) { st.loss += softplus_neg(r); st.correct += (x > 0.f) ? 1.f : 0.f; }
  float A = cf * c.plain_scale, Bc = c.kreg, Cc = 0.f;
  if (c.adver) {
    const float a = c.eps * cf * rsqrt_fast(fmaxf(cf * cf * s_dd, 1e-12f));
    const float b = c.eps * cf * rsqrt_fast(fmaxf(cf * cf * s_pp, 1e-12f));
    const float xa = x + 2.f * b * s_pp + a * s_dd + 2.f * a * b * x;
    float ra;
    const float ca = c.reg_adv * bpr_coeff(xa, ra);
    A = cf + ca;
    Bc = 2.f * ca * b + c.kreg;
    Cc = ca * a;
  }
  const float kreg = c.kreg, lr = c.lr;
#pragma unroll
  for (int k = 0; k < V; ++k) {
    const int e = (k * G + lane) * 4;
    if (FULL || e < d) {
      const float4 dq = f4_sub(q[k], n[k]);
      const float4 gp = f4_fma(A, dq, f4_scale(p[k], Bc));
      const float4 t = f4_fma(A, p[k], f4_scale(dq, Cc));
      const float4 gi = f4_fma(kreg, q[k], t);
      const float4 gj = f4_fma(kreg, n[k], f4_scale(t, -1.f));
      float4 a0 = ap[k], a1 = ai[k], a2 = aj[k], w0 = p[k], w1 = q[k], w2 = n[k];
#define APR_ADAGRAD_LANE(acc, w, g, f)            \
  acc.f = fmaf(g.f, g.f, acc.f);                  \
  w.f = fmaf(-lr * g.f, rsqrt_fast(acc.f), w.f);
      APR_ADAGRAD_LANE(a0, w0, gp, x) APR_ADAGRAD_LANE(a0, w0, gp, y) APR_ADAGRAD_LANE(a0, w0, gp, z) APR_ADAGRAD_LANE(a0, w0, gp, w)
      APR_ADAGRAD_LANE(a1, w1, gi, x) APR_ADAGRAD_LANE(a1, w1, gi, y) APR_ADAGRAD_LANE(a1, w1, gi, z) APR_ADAGRAD_LANE(a1, w1, gi, w)
      APR_ADAGRAD_LANE(a2, w2, gj, x) APR_ADAGRAD_LANE(a2, w2, gj, y) APR_ADAGRAD_LANE(a2, w2, gj, z) APR_ADAGRAD_LANE(a2, w2, gj, w)
#undef APR_ADAGRAD_LANE
      stcg4(Au + e, a0); stcg4(Pu + e, w0);
      stcg4(Ai + e, a1); stcg4(Qi + e, w1);
      stcg4(Aj + e, a2); stcg4(Qj + e, w2);
    }
  }
}

// fast segments k in [k0, k1) of the partitioned header array; one lane group per segment, grid-stride.
// Sharded tables: a segment is processed by the rank that OWNS ITS USER ROW (P and its accumulator stay local; only
// item rows cross NVLink).  Each lane fetches one header, a ballot finds the rank's own, shuffles broadcast them.
template <int G, int V, bool FULL>
__device__ __forceinline__ void fast_range(const StepCtx& c, int s, int k0, int k1, int gid, int ngroups, int lane,
                                           unsigned mask, StepStats& st) {
  const int4* seg_hdr = c.seg_hdr + int64_t(s) * c.B * 2;
  if (c.nranks == 1) {
    for (int k = k0 + gid; k < k1; k += ngroups) {
      const int user = __ldg(&seg_hdr[2 * k].x);
      const int2 ij = __ldg(reinterpret_cast<const int2*>(&seg_hdr[2 * k + 1]));
      fast_segment<G, V, FULL>(c, user, ij.x, ij.y, lane, mask, st);
    }
    return;
  }
  const int gshift = (threadIdx.x & 31) / G * G;  // first lane of this group inside its warp
  for (int base = k0 + gid * G; base < k1; base += ngroups * G) {
    const int k = base + lane;
    int user = -1;
    int2 ij = make_int2(0, 0);
    if (k < k1) {
      user = __ldg(&seg_hdr[2 * k].x);
      ij = __ldg(reinterpret_cast<const int2*>(&seg_hdr[2 * k + 1]));
    }
    const bool mine = k < k1 && (user & (c.nranks - 1)) == c.rank;
    unsigned m = (__ballot_sync(mask, mine) & mask) >> gshift;
    while (m) {
      const int src = __ffs(m) - 1;
      m &= m - 1;
      const int us = __shfl_sync(mask, user, src, G);
      const int is = __shfl_sync(mask, ij.x, src, G);
      const int js = __shfl_sync(mask, ij.y, src, G);
      fast_segment<G, V, FULL>(c, us, is, js, lane, mask, st);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// GENERAL segments (shared items and/or several triples of one user)
// ---------------------------------------------------------------------------------------------------------
// A complete step for a general segment that touches NO shared item (several triples of one user), or a complete
// BPR step for any segment (shared item gradients then go through item_sink's RED).
template <int G, int V>
__device__ __forceinline__ void segment_complete(const StepCtx& c, const int4 h0, const int4 h1, const int4* rec, int lane,
                                                 unsigned mask, StepStats& st) {
  const int d = c.d;
  const bool adver = c.adver != 0;
  const int user = h0.x;
  const int b0 = h0.y, b1 = h0.y + h0.z;
  float* Pu = P_ROW(c, user);
  float* Au = AP_ROW(c, user);
  Row<G, V> p, g;
  row_load<G, V>(p, Pu, lane, d);
  row_zero<G, V>(g);
  int4 rc = h1;
  for (int pos = b0; pos < b1; ++pos) {
    if (pos > b0) rc = rec[pos];
    Row<G, V> q, n;
    row_load<G, V>(q, Q_ROW(c, rc.x), lane, d);
    row_load<G, V>(n, Q_ROW(c, rc.y), lane, d);
    const float x = row_dot<G, V>(p, q, mask) - row_dot<G, V>(p, n, mask);
    float r;
    const float cf = bpr_coeff(x, r);
    if (st.on) { st.loss += softplus_neg(r); st.correct += (x > 0.f) ? 1.f : 0.f; }
    const float cs = adver ? cf : cf * c.plain_scale;
#pragma unroll
    for (int k = 0; k < V; ++k) g.v[k] = f4_fma(cs, f4_sub(q.v[k], n.v[k]), g.v[k]);
    if (!adver) {
      Row<G, V> hi, hj;
#pragma unroll
      for (int k = 0; k < V; ++k) {
        g.v[k] = f4_fma(c.kreg, p.v[k], g.v[k]);
        hi.v[k] = f4_fma(c.kreg, q.v[k], f4_scale(p.v[k], cs));
        hj.v[k] = f4_fma(c.kreg, n.v[k], f4_scale(p.v[k], -cs));
      }
      item_sink<G, V>(c, rc.x, rc.z, q, hi, lane, d);
      item_sink<G, V>(c, rc.y, rc.w, n, hj, lane, d);
    }
  }
  if (adver) {
    // second pass right away: Delta from the local gradients (x, c recomputed: same bits)
    const float sp = delta_scale<G, V>(g, c.eps, mask);
    Row<G, V> pd;
#pragma unroll
    for (int k = 0; k < V; ++k) pd.v[k] = f4_fma(sp, g.v[k], p.v[k]);
    rc = h1;
    for (int pos = b0; pos < b1; ++pos) {
      if (pos > b0) rc = rec[pos];
      Row<G, V> q, n;
      row_load<G, V>(q, Q_ROW(c, rc.x), lane, d);
      row_load<G, V>(n, Q_ROW(c, rc.y), lane, d);
      const float x = row_dot<G, V>(p, q, mask) - row_dot<G, V>(p, n, mask);
      float r;
      const float cf = bpr_coeff(x, r);
      adv_triple<G, V>(c, rc, cf, p, pd, q, n, g, lane, d, mask);
    }
  }
  adagrad_row<G, V>(Pu, Au, p, g, lane, d, c.lr);
}

// APR stage 0 for one segment that touches a shared item: plain forward/backward; publishes G_P[k], c per triple, G_Q.
template <int G, int V>
__device__ __forceinline__ void slow_plain(const StepCtx& c, int k0, const int4 h0, const int4 h1, const int4* rec, int lane,
                                           unsigned mask, StepStats& st) {
  const int d = c.d;
  const int b0 = h0.y, b1 = h0.y + h0.z;
  Row<G, V> p, g;
  row_load<G, V>(p, P_ROW(c, h0.x), lane, d);
  row_zero<G, V>(g);
  int4 rc = h1;
  for (int pos = b0; pos < b1; ++pos) {
    if (pos > b0) rc = rec[pos];
    Row<G, V> q, n;
    row_load<G, V>(q, Q_ROW(c, rc.x), lane, d);
    row_load<G, V>(n, Q_ROW(c, rc.y), lane, d);
    const float x = row_dot<G, V>(p, q, mask) - row_dot<G, V>(p, n, mask);
    float r;
    const float cf = bpr_coeff(x, r);
    if (st.on) { st.loss += softplus_neg(r); st.correct += (x > 0.f) ? 1.f : 0.f; }
    Row<G, V> t;
#pragma unroll
    for (int k = 0; k < V; ++k) { g.v[k] = f4_fma(cf, f4_sub(q.v[k], n.v[k]), g.v[k]); t.v[k] = f4_scale(p.v[k], cf); }
    if (rc.z >= 0) row_red<G, V>(t, GQ_ROW(c, rc.z), lane, d);
    if (rc.w >= 0) {
#pragma unroll
      for (int k = 0; k < V; ++k) t.v[k] = f4_scale(t.v[k], -1.f);
      row_red<G, V>(t, GQ_ROW(c, rc.w), lane, d);
    }
    if (lane == 0) c.cbuf[pos] = cf;
    if (c.nranks > 1) {
      // nobody writes these two rows between stage 0 and stage 1 of a step (a shared item row changes in stage 2 only, a
      // singleton item row only in this segment's own stage 1): stage 1 takes them from the local copy
      row_store<G, V>(q, c.QN + (int64_t(pos) * 2) * d, lane, d);
      row_store<G, V>(n, c.QN + (int64_t(pos) * 2 + 1) * d, lane, d);
    }
  }
  row_store<G, V>(g, c.GP + int64_t(k0) * d, lane, d);
}

// APR stage 1 for one segment that touches a shared item: Delta from the batch-wide sums, adversarial pass, Adagrad on
// the rows this segment owns (user row, singleton item rows); shared item gradients are RED into H_Q.
template <int G, int V>
__device__ __forceinline__ void slow_adv(const StepCtx& c, int k0, const int4 h0, const int4 h1, const int4* rec, int lane,
                                         unsigned mask) {
  const int d = c.d;
  const int b0 = h0.y, b1 = h0.y + h0.z;
  int4 rc = h1;
  Row<G, V> p, g, pd;
  row_load<G, V>(p, P_ROW(c, h0.x), lane, d);
  row_load<G, V>(g, c.GP + int64_t(k0) * d, lane, d);
  const float sp = delta_scale<G, V>(g, c.eps, mask);
#pragma unroll
  for (int k = 0; k < V; ++k) pd.v[k] = f4_fma(sp, g.v[k], p.v[k]);
  for (int pos = b0; pos < b1; ++pos) {
    if (pos > b0) rc = rec[pos];
    const float cf = __ldcg(&c.cbuf[pos]);
    Row<G, V> q, n;
    row_load<G, V>(q, c.nranks > 1 ? c.QN + (int64_t(pos) * 2) * d : Q_ROW(c, rc.x), lane, d);
    row_load<G, V>(n, c.nranks > 1 ? c.QN + (int64_t(pos) * 2 + 1) * d : Q_ROW(c, rc.y), lane, d);
    adv_triple<G, V>(c, rc, cf, p, pd, q, n, g, lane, d, mask);
  }
  adagrad_row<G, V>(P_ROW(c, h0.x), AP_ROW(c, h0.x), p, g, lane, d, c.lr);
}

// stage 2: shared item slots
template <int G, int V>
__device__ __forceinline__ void items_shared(const StepCtx& c, int s, int gid, int ngroups, int lane) {
  const int d = c.d, B = c.B;
  const int ni = c.icnt[s];
  const int32_t* iu_item = c.iu_item + int64_t(s) * B;
  Row<G, V> z;
  row_zero<G, V>(z);
  for (int slot = gid * c.nranks + c.rank; slot < ni; slot += ngroups * c.nranks) {
    const int item = iu_item[slot];
    if (item < 0) continue;   // slot of a pair work unit: updated there
    Row<G, V> w, g, a;
    row_load<G, V>(g, HQ_ROW(c, slot), lane, d);
    row_load<G, V>(w, Q_ROW(c, item), lane, d);
    row_load<G, V>(a, AQ_ROW(c, item), lane, d);
    adagrad_apply<G, V>(Q_ROW(c, item), AQ_ROW(c, item), w, a, g, lane, d, c.lr);
    row_store<G, V>(z, HQ_ROW(c, slot), lane, d);
    if (c.adver) row_store<G, V>(z, GQ_ROW(c, slot), lane, d);
  }
}

// ---------------------------------------------------------------------------------------------------------
// PAIR work unit: two single-triple segments a, b whose only coupling to the rest of the batch is ONE item that occurs
// in both (and nowhere else).  Everything the three-stage general path does through the workspace -- the batch-wide
// gradient sum of the shared item, its Delta, its total gradient -- is a two-term sum here, so one group does both
// triples in registers: 6 row reads (the shared row twice), 5 accumulator reads, 10 row writes.
// Same formulas as slow_plain / adv_triple / item_sink (APR.py:143-165,180-195), per occurrence.
template <int G, int V>
__device__ __forceinline__ void pair_unit(const StepCtx& c, const int4 ph, const int4 ra, const int4 rb, int lane,
                                          unsigned mask, StepStats& st) {
  const int d = c.d;
  const int slot = ph.z;
  const bool sa = ra.z == slot, sb = rb.z == slot;      // shared item is the POSITIVE of a / of b (else the negative)
  float* Pa = P_ROW(c, ph.x);
  float* Pb = P_ROW(c, ph.y);
  float* Qa = Q_ROW(c, ra.x);
  float* Na = Q_ROW(c, ra.y);
  float* Qb = Q_ROW(c, rb.x);
  float* Nb = Q_ROW(c, rb.y);
  const int item_s = sa ? ra.x : ra.y, item_oa = sa ? ra.y : ra.x, item_ob = sb ? rb.y : rb.x;
  Row<G, V> pa, pb, qa, na, qb, nb, aPa, aPb, aS, aOa, aOb;
  row_load<G, V>(pa, Pa, lane, d);
  row_load<G, V>(pb, Pb, lane, d);
  row_load<G, V>(qa, Qa, lane, d);
  row_load<G, V>(na, Na, lane, d);
  row_load<G, V>(qb, Qb, lane, d);
  row_load<G, V>(nb, Nb, lane, d);
  row_load<G, V>(aPa, AP_ROW(c, ph.x), lane, d);
  row_load<G, V>(aPb, AP_ROW(c, ph.y), lane, d);
  row_load<G, V>(aS, AQ_ROW(c, item_s), lane, d);
  row_load<G, V>(aOa, AQ_ROW(c, item_oa), lane, d);
  row_load<G, V>(aOb, AQ_ROW(c, item_ob), lane, d);

  const float xa = row_dot<G, V>(pa, qa, mask) - row_dot<G, V>(pa, na, mask);
  const float xb = row_dot<G, V>(pb, qb, mask) - row_dot<G, V>(pb, nb, mask);
  float r0, r1;
  const float ps = c.adver ? 1.0f : c.plain_scale;
  const float ca = ps * bpr_coeff(xa, r0), cb = ps * bpr_coeff(xb, r1);
  if (st.on) {
    st.loss += softplus_neg(r0) + softplus_neg(r1);
    st.correct += ((xa > 0.f) ? 1.f : 0.f) + ((xb > 0.f) ? 1.f : 0.f);
  }
  Row<G, V> ga, gb, hia, hja, hib, hjb;   // user-row gradients; per-occurrence item gradients (pos / neg of a and b)
#pragma unroll
  for (int k = 0; k < V; ++k) {
    ga.v[k] = f4_scale(f4_sub(qa.v[k], na.v[k]), ca);
    gb.v[k] = f4_scale(f4_sub(qb.v[k], nb.v[k]), cb);
  }
  if (!c.adver) {
#pragma unroll
    for (int k = 0; k < V; ++k) {
      ga.v[k] = f4_fma(c.kreg, pa.v[k], ga.v[k]);
      gb.v[k] = f4_fma(c.kreg, pb.v[k], gb.v[k]);
      hia.v[k] = f4_fma(c.kreg, qa.v[k], f4_scale(pa.v[k], ca));
      hja.v[k] = f4_fma(c.kreg, na.v[k], f4_scale(pa.v[k], -ca));
      hib.v[k] = f4_fma(c.kreg, qb.v[k], f4_scale(pb.v[k], cb));
      hjb.v[k] = f4_fma(c.kreg, nb.v[k], f4_scale(pb.v[k], -cb));
    }
  } else {
    // plain gradient sum of the shared item over its two occurrences, and the three Delta scales that involve sums
    Row<G, V> gs;
#pragma unroll
    for (int k = 0; k < V; ++k) gs.v[k] = f4_fma(sb ? cb : -cb, pb.v[k], f4_scale(pa.v[k], sa ? ca : -ca));
    const float scs = delta_scale<G, V>(gs, c.eps, mask);
    const float spa = delta_scale<G, V>(ga, c.eps, mask), spb = delta_scale<G, V>(gb, c.eps, mask);
    // singleton occurrences: G = +-c p, so Delta = eps c p rsqrt(max(c^2 |p|^2, 1e-12))
    Row<G, V> t;
#pragma unroll
    for (int k = 0; k < V; ++k) t.v[k] = f4_scale(pa.v[k], ca);
    const float soa = delta_scale<G, V>(t, c.eps, mask);
#pragma unroll
    for (int k = 0; k < V; ++k) t.v[k] = f4_scale(pb.v[k], cb);
    const float sob = delta_scale<G, V>(t, c.eps, mask);
    Row<G, V> pda, pdb, qda, nda, qdb, ndb;
#pragma unroll
    for (int k = 0; k < V; ++k) {
      pda.v[k] = f4_fma(spa, ga.v[k], pa.v[k]);
      pdb.v[k] = f4_fma(spb, gb.v[k], pb.v[k]);
      // Delta of an occurrence: the shared item's from gs, a singleton's from its own +-c p
      qda.v[k] = sa ? f4_fma(scs, gs.v[k], qa.v[k]) : f4_fma(soa * ca, pa.v[k], qa.v[k]);
      nda.v[k] = sa ? f4_fma(-soa * ca, pa.v[k], na.v[k]) : f4_fma(scs, gs.v[k], na.v[k]);
      qdb.v[k] = sb ? f4_fma(scs, gs.v[k], qb.v[k]) : f4_fma(sob * cb, pb.v[k], qb.v[k]);
      ndb.v[k] = sb ? f4_fma(-sob * cb, pb.v[k], nb.v[k]) : f4_fma(scs, gs.v[k], nb.v[k]);
    }
    const float xaa = row_dot<G, V>(pda, qda, mask) - row_dot<G, V>(pda, nda, mask);
    const float xab = row_dot<G, V>(pdb, qdb, mask) - row_dot<G, V>(pdb, ndb, mask);
    float r2, r3;
    const float caa = c.reg_adv * bpr_coeff(xaa, r2), cab = c.reg_adv * bpr_coeff(xab, r3);
#pragma unroll
    for (int k = 0; k < V; ++k) {
      ga.v[k] = f4_fma(caa, f4_sub(qda.v[k], nda.v[k]), ga.v[k]);
      ga.v[k] = f4_fma(c.kreg, pa.v[k], ga.v[k]);
      gb.v[k] = f4_fma(cab, f4_sub(qdb.v[k], ndb.v[k]), gb.v[k]);
      gb.v[k] = f4_fma(c.kreg, pb.v[k], gb.v[k]);
      const float4 ta = f4_fma(caa, pda.v[k], f4_scale(pa.v[k], ca));   // c p + reg_adv c' (p + dP)
      const float4 tb = f4_fma(cab, pdb.v[k], f4_scale(pb.v[k], cb));
      hia.v[k] = f4_fma(c.kreg, qa.v[k], ta);
      hja.v[k] = f4_fma(c.kreg, na.v[k], f4_scale(ta, -1.f));
      hib.v[k] = f4_fma(c.kreg, qb.v[k], tb);
      hjb.v[k] = f4_fma(c.kreg, nb.v[k], f4_scale(tb, -1.f));
    }
  }
  // total gradient of the shared item = its two occurrences; the other item of each triple keeps its own
  Row<G, V> hs;
#pragma unroll
  for (int k = 0; k < V; ++k) {
    const float4 ua = sa ? hia.v[k] : hja.v[k], ub = sb ? hib.v[k] : hjb.v[k];
    hs.v[k] = make_float4(ua.x + ub.x, ua.y + ub.y, ua.z + ub.z, ua.w + ub.w);
  }
  adagrad_apply<G, V>(Pa, AP_ROW(c, ph.x), pa, aPa, ga, lane, d, c.lr);
  adagrad_apply<G, V>(Pb, AP_ROW(c, ph.y), pb, aPb, gb, lane, d, c.lr);
  adagrad_apply<G, V>(sa ? Na : Qa, AQ_ROW(c, item_oa), sa ? na : qa, aOa, sa ? hja : hia, lane, d, c.lr);
  adagrad_apply<G, V>(sb ? Nb : Qb, AQ_ROW(c, item_ob), sb ? nb : qb, aOb, sb ? hjb : hib, lane, d, c.lr);
  adagrad_apply<G, V>(sa ? Qa : Na, AQ_ROW(c, item_s), sa ? qa : na, aS, hs, lane, d, c.lr);
}

template <int G, int V>
__device__ __forceinline__ void pair_range(const StepCtx& c, int s, int gid, int ngroups, int lane, unsigned mask,
                                           StepStats& st) {
  if (!c.npair) return;
  const int np = c.npair[s];
  const int4* pairs = c.pairs + int64_t(s) * (c.B / 2 + 1) * 3;
  for (int k = gid; k < np; k += ngroups) {
    const int4 ph = __ldg(&pairs[3 * k]);
    if (c.nranks > 1 && (ph.x & (c.nranks - 1)) != c.rank) continue;   // sharded tables: the owner of user a does the pair
    const int4 ra = __ldg(&pairs[3 * k + 1]), rb = __ldg(&pairs[3 * k + 2]);
    pair_unit<G, V>(c, ph, ra, rb, lane, mask, st);
  }
}

// The stages of the GENERAL path for one step (grid-wide ordering between them):
//   stage 0 (APR)  segments with a shared item: slow_plain
//   stage 1        APR: slow_adv for those, segment_complete for the rest;  BPR: segment_complete for all
//   stage 2        items_shared
template <int G, int V>
__device__ __forceinline__ void general_stage(const StepCtx& c, int s, int stage, int gid, int ngroups, int lane,
                                              unsigned mask, StepStats& st) {
  const int B = c.B;
  if (stage == 2) { items_shared<G, V>(c, s, gid, ngroups, lane); return; }
  const int ng = c.nslow[s];
  const int4* seg_hdr = c.seg_hdr + int64_t(s) * B * 2;
  const int4* rec = c.rec + int64_t(s) * B;
  const int gshift = (threadIdx.x & 31) / G * G;
  const int C = min(G, c.nranks);  // headers fetched per iteration: ~one of them belongs to this rank
  for (int base = gid * C; base < ng; base += ngroups * C) {
   // one header per lane, ballot the segments whose user row this rank owns, process them one by one
   int4 hv0 = make_int4(-1, 0, 0, 0), hv1 = make_int4(0, 0, 0, 0);
   const bool have = lane < C && base + lane < ng;
   if (have) { hv0 = __ldg(&seg_hdr[2 * (base + lane)]); hv1 = __ldg(&seg_hdr[2 * (base + lane) + 1]); }
   const bool mine = have && (hv0.x & (c.nranks - 1)) == c.rank;
   unsigned m = (__ballot_sync(mask, mine) & mask) >> gshift;
   while (m) {
    const int src = __ffs(m) - 1;
    m &= m - 1;
    const int k0 = base + src;
    const int4 h0 = shfl4<G>(mask, hv0, src);
    const int4 h1 = shfl4<G>(mask, hv1, src);
    const bool shared = h0.w != 0;
    if (stage == 0) {
      if (shared) slow_plain<G, V>(c, k0, h0, h1, rec, lane, mask, st);
    } else if (shared && c.adver) {
      slow_adv<G, V>(c, k0, h0, h1, rec, lane, mask);
    } else {
      segment_complete<G, V>(c, h0, h1, rec, lane, mask, st);
    }
   }
  }
}

constexpr int kThreads = 256;

template <int G>
__device__ __forceinline__ unsigned group_mask() {
  const unsigned wl = threadIdx.x & 31;
  return (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (wl / G * G));
}

// mode 0: general-path stage kernel (small grid) ...
// step / stats of a mode-0 kernel: straight from the parameters, or relative to the device-side cursor when the launch
// is a node of a replayed CUDA graph
#define APR_STEP_ARGS(c, s_param)                                   \
  const int s = (c).dyn ? (c).dyn->s + (s_param) : (s_param);       \
  float* const stats = (c).dyn ? (c).dyn->stats : (c).stats;        \
  StepStats st = {0.f, 0.f, stats != nullptr};

template <int G, int V>
__global__ void __launch_bounds__(kThreads, (V == 1 ? 4 : 1)) general_stage_kernel(StepCtx c, int s_param, int stage) {
  const int lane = threadIdx.x % G;
  const int gid = (blockIdx.x * kThreads + threadIdx.x) / G;
  const int ngroups = gridDim.x * (kThreads / G);
  if (c.abort && *c.abort) return;
  APR_STEP_ARGS(c, s_param)
  if (stage == 0 && !(c.adver && c.icnt[s] > 0)) return;
  general_stage<G, V>(c, s, stage, gid, ngroups, lane, group_mask<G>(), st);
  if (stats && stage < 2) stats_flush(stats, s, st.loss, st.correct, lane == 0);
}

// ... and the fast-path kernel, launched on a second stream so that the general path's load chains hide under it
template <int G, int V, bool FULL>
__global__ void __launch_bounds__(kThreads, (V == 1 ? 4 : (V == 2 ? 2 : 1))) fast_kernel(StepCtx c, int s_param) {
  const int lane = threadIdx.x % G;
  const int gid = (blockIdx.x * kThreads + threadIdx.x) / G;
  const int ngroups = gridDim.x * (kThreads / G);
  if (c.abort && *c.abort) return;
  APR_STEP_ARGS(c, s_param)
  fast_range<G, V, FULL>(c, s, c.nslow[s] + (c.npair ? 2 * c.npair[s] : 0), c.ucnt[s], gid, ngroups, lane, group_mask<G>(), st);
  if (stats) stats_flush(stats, s, st.loss, st.correct, lane == 0);
}

// mode 0: the pair work units of one step (third stream; touches rows no other kernel of the step touches)
template <int G, int V>
__global__ void __launch_bounds__(kThreads, 2) pair_kernel(StepCtx c, int s_param) {
  const int lane = threadIdx.x % G;
  const int gid = (blockIdx.x * blockDim.x + threadIdx.x) / G;      // launched with 256 or 128 threads per block
  const int ngroups = gridDim.x * (blockDim.x / G);
  if (c.abort && *c.abort) return;
  APR_STEP_ARGS(c, s_param)
  pair_range<G, V>(c, s, gid, ngroups, lane, group_mask<G>(), st);
  if (stats) stats_flush(stats, s, st.loss, st.correct, lane == 0);
}

// ---------------------------------------------------------------------------------------------------------
// --adv random (APR.py:170-177; shape-consistent form evaluation_adv.py:182-189): at every step
//     Delta_P = eps * l2_normalize(truncated_normal([rows, d], 0, 0.01)),  Delta_Q likewise,
// drawn afresh for ALL rows; only the rows a batch touches matter, so the noise of a row is generated where the row is
// used -- counter-based (Philox: element = row * d + column, table id = 2 * global_step + {0: P, 1: Q}, key = (seed,
// STREAM_ADV)), never a table.  Delta does not depend on the batch, so there is no plain stage: one kernel does every
// segment and pair (forward, perturbed forward, total gradient, Adagrad; shared items RED into H_Q), then stage 2.
// ---------------------------------------------------------------------------------------------------------
template <int G, int V>
__device__ __forceinline__ void noise_delta(Row<G, V>& dl, const StepCtx& c, uint32_t table_id, int row, int lane,
                                            unsigned mask) {
  const int d = c.d;
#pragma unroll
  for (int k = 0; k < V; ++k) {
    const int e = (k * G + lane) * 4;
    float4 z = f4_zero();
    if (e < d) {
      const uint64_t base = uint64_t(row) * uint64_t(d) + uint64_t(e);
      z.x = truncated_normal_elem(base, c.noise_seed, table_id, kStreamAdv) * 0.01f;
      z.y = truncated_normal_elem(base + 1, c.noise_seed, table_id, kStreamAdv) * 0.01f;
      z.z = truncated_normal_elem(base + 2, c.noise_seed, table_id, kStreamAdv) * 0.01f;
      z.w = truncated_normal_elem(base + 3, c.noise_seed, table_id, kStreamAdv) * 0.01f;
    }
    dl.v[k] = z;
  }
  const float sc = delta_scale<G, V>(dl, c.eps, mask);
#pragma unroll
  for (int k = 0; k < V; ++k) dl.v[k] = f4_scale(dl.v[k], sc);
}

// one triple: plain + perturbed forward, adds the user-row gradient into g, returns the item-row gradients hi / hj
template <int G, int V>
__device__ __forceinline__ void random_triple(const StepCtx& c, uint32_t tq, const int4 rc, const Row<G, V>& p,
                                              const Row<G, V>& pd, const Row<G, V>& q, const Row<G, V>& n, Row<G, V>& g,
                                              Row<G, V>& hi, Row<G, V>& hj, int lane, unsigned mask, StepStats& st) {
  const float x = row_dot<G, V>(p, q, mask) - row_dot<G, V>(p, n, mask);
  float r;
  const float cf = bpr_coeff(x, r);
  if (st.on) { st.loss += softplus_neg(r); st.correct += (x > 0.f) ? 1.f : 0.f; }
  Row<G, V> qd, nd;
  noise_delta<G, V>(qd, c, tq, rc.x, lane, mask);
  noise_delta<G, V>(nd, c, tq, rc.y, lane, mask);
#pragma unroll
  for (int k = 0; k < V; ++k) { qd.v[k] = f4_add(q.v[k], qd.v[k]); nd.v[k] = f4_add(n.v[k], nd.v[k]); }
  const float xa = row_dot<G, V>(pd, qd, mask) - row_dot<G, V>(pd, nd, mask);
  float ra;
  const float ca = c.reg_adv * bpr_coeff(xa, ra);
#pragma unroll
  for (int k = 0; k < V; ++k) {
    g.v[k] = f4_fma(cf, f4_sub(q.v[k], n.v[k]), g.v[k]);
    g.v[k] = f4_fma(ca, f4_sub(qd.v[k], nd.v[k]), g.v[k]);
    g.v[k] = f4_fma(c.kreg, p.v[k], g.v[k]);
    const float4 t = f4_fma(ca, pd.v[k], f4_scale(p.v[k], cf));  // c p + reg_adv c' (p + dP)
    hi.v[k] = f4_fma(c.kreg, q.v[k], t);
    hj.v[k] = f4_fma(c.kreg, n.v[k], f4_scale(t, -1.f));
  }
}

template <int G, int V>
__device__ __forceinline__ void random_segment(const StepCtx& c, uint32_t tp, const int4 h0, const int4 h1, const int4* rec,
                                               int lane, unsigned mask, StepStats& st) {
  const int d = c.d;
  float* Pu = P_ROW(c, h0.x);
  Row<G, V> p, pd, g;
  row_load<G, V>(p, Pu, lane, d);
  noise_delta<G, V>(pd, c, tp, h0.x, lane, mask);
#pragma unroll
  for (int k = 0; k < V; ++k) pd.v[k] = f4_add(p.v[k], pd.v[k]);
  row_zero<G, V>(g);
  int4 rc = h1;
  for (int pos = h0.y; pos < h0.y + h0.z; ++pos) {
    if (pos > h0.y) rc = rec[pos];
    Row<G, V> q, n, hi, hj;
    row_load<G, V>(q, Q_ROW(c, rc.x), lane, d);
    row_load<G, V>(n, Q_ROW(c, rc.y), lane, d);
    random_triple<G, V>(c, tp + 1u, rc, p, pd, q, n, g, hi, hj, lane, mask, st);
    item_sink<G, V>(c, rc.x, rc.z, q, hi, lane, d);
    item_sink<G, V>(c, rc.y, rc.w, n, hj, lane, d);
  }
  adagrad_row<G, V>(Pu, AP_ROW(c, h0.x), p, g, lane, d, c.lr);
}

// pair work unit (two single-triple segments coupled by one item that occurs in both and nowhere else; its workspace
// slot is struck from stage 2): the shared item's total gradient is the sum of its two occurrences, in registers
template <int G, int V>
__device__ __forceinline__ void random_pair(const StepCtx& c, uint32_t tp, const int4 ph, const int4 ra, const int4 rb,
                                            int lane, unsigned mask, StepStats& st) {
  const int d = c.d;
  const bool sa = ra.z == ph.z, sb = rb.z == ph.z;   // the shared item is the POSITIVE of a / of b (else the negative)
  Row<G, V> pa, pb, qa, na, qb, nb, pda, pdb, ga, gb, hia, hja, hib, hjb;
  row_load<G, V>(pa, P_ROW(c, ph.x), lane, d);
  row_load<G, V>(pb, P_ROW(c, ph.y), lane, d);
  row_load<G, V>(qa, Q_ROW(c, ra.x), lane, d);
  row_load<G, V>(na, Q_ROW(c, ra.y), lane, d);
  row_load<G, V>(qb, Q_ROW(c, rb.x), lane, d);
  row_load<G, V>(nb, Q_ROW(c, rb.y), lane, d);
  noise_delta<G, V>(pda, c, tp, ph.x, lane, mask);
  noise_delta<G, V>(pdb, c, tp, ph.y, lane, mask);
#pragma unroll
  for (int k = 0; k < V; ++k) { pda.v[k] = f4_add(pa.v[k], pda.v[k]); pdb.v[k] = f4_add(pb.v[k], pdb.v[k]); }
  row_zero<G, V>(ga);
  row_zero<G, V>(gb);
  random_triple<G, V>(c, tp + 1u, ra, pa, pda, qa, na, ga, hia, hja, lane, mask, st);
  random_triple<G, V>(c, tp + 1u, rb, pb, pdb, qb, nb, gb, hib, hjb, lane, mask, st);
  Row<G, V> hs;
#pragma unroll
  for (int k = 0; k < V; ++k) hs.v[k] = f4_add(sa ? hia.v[k] : hja.v[k], sb ? hib.v[k] : hjb.v[k]);
  adagrad_row<G, V>(P_ROW(c, ph.x), AP_ROW(c, ph.x), pa, ga, lane, d, c.lr);
  adagrad_row<G, V>(P_ROW(c, ph.y), AP_ROW(c, ph.y), pb, gb, lane, d, c.lr);
  const int item_s = sa ? ra.x : ra.y, item_oa = sa ? ra.y : ra.x, item_ob = sb ? rb.y : rb.x;
  adagrad_row<G, V>(Q_ROW(c, item_oa), AQ_ROW(c, item_oa), sa ? na : qa, sa ? hja : hia, lane, d, c.lr);
  adagrad_row<G, V>(Q_ROW(c, item_ob), AQ_ROW(c, item_ob), sb ? nb : qb, sb ? hjb : hib, lane, d, c.lr);
  adagrad_row<G, V>(Q_ROW(c, item_s), AQ_ROW(c, item_s), sa ? qa : na, hs, lane, d, c.lr);
}

template <int G, int V>
__global__ void __launch_bounds__(kThreads) random_step_kernel(StepCtx c, int s) {
  const int lane = threadIdx.x % G;
  const int gid = (blockIdx.x * kThreads + threadIdx.x) / G;
  const int ngroups = gridDim.x * (kThreads / G);
  const unsigned mask = group_mask<G>();
  StepStats st = {0.f, 0.f, c.stats != nullptr};
  const uint32_t tp = (c.noise_step0 + uint32_t(s)) << 1;          // table id of Delta_P at this step; Delta_Q: tp + 1
  const int ng = c.nslow[s], np = c.npair ? c.npair[s] : 0, nu = c.ucnt[s];
  const int first_fast = ng + 2 * np;
  const int n_seg = ng + (nu - first_fast);
  const int4* seg_hdr = c.seg_hdr + int64_t(s) * c.B * 2;
  const int4* rec = c.rec + int64_t(s) * c.B;
  const int4* pairs = c.pairs + int64_t(s) * (c.B / 2 + 1) * 3;
  for (int w = gid; w < n_seg + np; w += ngroups) {
    if (w < n_seg) {
      const int k = w < ng ? w : first_fast + (w - ng);
      random_segment<G, V>(c, tp, __ldg(&seg_hdr[2 * k]), __ldg(&seg_hdr[2 * k + 1]), rec, lane, mask, st);
    } else {
      if constexpr (V == 1) {
        const int k = w - n_seg;
        random_pair<G, V>(c, tp, __ldg(&pairs[3 * k]), __ldg(&pairs[3 * k + 1]), __ldg(&pairs[3 * k + 2]), lane, mask, st);
      }
    }
  }
  if (c.stats) stats_flush(c.stats, s, st.loss, st.correct, lane == 0);
}

template <int G, int V>
static int run_random_steps(const StepCtx& c, cudaStream_t st) {
  const int gpb = kThreads / G;
  const int grid = std::max(1, std::min((c.B + gpb - 1) / gpb, sm_count() * 4));
  const int grid_gen = std::max(1, std::min((c.B + gpb - 1) / gpb, sm_count() * 2));
  for (int s = c.s_begin; s < c.s_end; ++s) {
    random_step_kernel<G, V><<<grid, kThreads, 0, st>>>(c, s);
    general_stage_kernel<G, V><<<grid_gen, kThreads, 0, st>>>(c, s, 2);   // Adagrad on the shared item rows
  }
  APR_LAUNCH_CHECK();
  return APR_OK;
}

static int dispatch_random_steps(const StepCtx& c, cudaStream_t st) {
  const int q = c.d / 4;
  if (q <= 4) return run_random_steps<4, 1>(c, st);
  if (q <= 8) return run_random_steps<8, 1>(c, st);
  if (q <= 16) return run_random_steps<16, 1>(c, st);
  if (q <= 32) return run_random_steps<32, 1>(c, st);
  if (q <= 64) return run_random_steps<32, 2>(c, st);
  if (q <= 96) return run_random_steps<32, 3>(c, st);
  return run_random_steps<32, 4>(c, st);
}

// device-side step cursor of the replayed graphs (StepDyn): set before the first graph launch of a call, advanced by the
// last node of every graph
__global__ void dyn_set_kernel(StepDyn* dyn, int s, float* stats) { dyn->s = s; dyn->stats = stats; }
__global__ void dyn_advance_kernel(StepDyn* dyn, int n) { dyn->s += n; }

// mode 1: all steps in one cooperative launch; the general stages run under the two halves of the fast range
template <int G, int V, bool FULL>
__global__ void __launch_bounds__(kThreads) step_persistent_kernel(StepCtx c) {
  cg::grid_group grid = cg::this_grid();
  // mode 2: the whole grid is one cluster (<= 16 CTAs) -> hardware cluster barrier (release/acquire at cluster scope)
  // instead of the cooperative grid barrier; mode 1: cooperative launch over the whole GPU
  auto sync_all = [&]() { if (c.cluster_sync) cg::this_cluster().sync(); else grid.sync(); };
  const int lane = threadIdx.x % G;
  const int gid = (blockIdx.x * kThreads + threadIdx.x) / G;
  const int ngroups = gridDim.x * (kThreads / G);
  const unsigned mask = group_mask<G>();
  for (int s = c.s_begin; s < c.s_end; ++s) {
    const int nu = c.ucnt[s], ng = c.nslow[s] + (c.npair ? 2 * c.npair[s] : 0);   // ng: first fast segment
    const bool shared = c.icnt[s] > 0;  // grid-uniform
    const int mid = ng + (nu - ng + 1) / 2;
    StepStats st = {0.f, 0.f, c.stats != nullptr};
    if (shared && c.adver) {
      general_stage<G, V>(c, s, 0, gid, ngroups, lane, mask, st);
      sync_all();
    }
    general_stage<G, V>(c, s, 1, gid, ngroups, lane, mask, st);
    if constexpr (V == 1) pair_range<G, V>(c, s, gid, ngroups, lane, mask, st);   // independent of every other segment
    fast_range<G, V, FULL>(c, s, ng, mid, gid, ngroups, lane, mask, st);
    if (shared) {
      sync_all();  // every H_Q contribution must have landed before the shared rows are updated
      general_stage<G, V>(c, s, 2, gid, ngroups, lane, mask, st);
    }
    fast_range<G, V, FULL>(c, s, mid, nu, gid, ngroups, lane, mask, st);
    if (c.stats) stats_flush(c.stats, s, st.loss, st.correct, lane == 0);
    sync_all();
  }
}

static int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}

// ------------------------------------------------------------------------------------------------
// CUDA-graph replay of mode 0.  The launches of a GROUP of n steps (n in kGraphGroups: per step the fast kernel, the pair
// kernel, the general stages and the fork/join edges of the three streams, then one node that advances the device-side
// step cursor) are captured ONCE into an executable graph whose kernel parameters do not depend on which steps it runs:
// the step index is `dyn->s + k` with dyn in the workspace header.  Executable graphs are cached per device, keyed by
// the bytes of the static StepCtx (tables, workspace, hyper-parameters, batch, d) + kernel instantiation + n; when a key
// is seen for the first time ALL group sizes are built, so no cudaGraphInstantiate can happen later for any step count
// of that configuration (round 1 keyed a 4-slot ring by the call's step count: a 20-step call after a 5-step warm-up
// re-instantiated inside the timed region).  A full cache recycles its least recently used graph of the same topology
// through cudaGraphExecUpdate (parameters only).
// ------------------------------------------------------------------------------------------------
constexpr int kGraphGroups[] = {32, 16, 8, 4, 2, 1};
constexpr size_t kGraphCacheMax = 96;

static GraphEntry* graph_find(DeviceContext& ax, const StepCtx& key, const void* fn, int n) {
  for (auto& g : ax.graphs)
    if (g.exec && g.fn == fn && g.n_steps == n && g.key.size() == sizeof(StepCtx) && memcmp(g.key.data(), &key, sizeof(StepCtx)) == 0)
      return &g;
  return nullptr;
}

// capture `issue(capture_stream, n)` and turn it into the executable graph of (key, fn, n)
template <typename Issue>
static int graph_build(DeviceContext& ax, const StepCtx& key, const void* fn, int n, int topo, Issue&& issue, GraphEntry** out) {
  APR_CUDA_CHECK(cudaStreamBeginCapture(ax.capture_stream, cudaStreamCaptureModeThreadLocal));
  const int rc = issue(ax.capture_stream, n);
  cudaGraph_t graph = nullptr;
  const cudaError_t ee = cudaStreamEndCapture(ax.capture_stream, &graph);
  if (rc != APR_OK || ee != cudaSuccess || !graph) {
    if (graph) cudaGraphDestroy(graph);
    if (ee != cudaSuccess) { set_cuda_error(ee, "cudaStreamEndCapture"); return APR_E_CUDA; }
    return rc != APR_OK ? rc : APR_E_CUDA;
  }
  GraphEntry* e = nullptr;
  if (ax.graphs.size() >= kGraphCacheMax) {
    // recycle: least recently used entry, preferring one whose topology matches (update in place, no instantiate)
    for (int pass = 0; pass < 2 && !e; ++pass)
      for (auto& g : ax.graphs)
        if ((pass == 1 || (g.fn == fn && g.n_steps == n && g.topo == topo)) && (!e || g.last_use < e->last_use)) e = &g;
  } else {
    ax.graphs.reserve(kGraphCacheMax);   // entries are handed out by pointer: never reallocate
    ax.graphs.emplace_back();
    e = &ax.graphs.back();
  }
  cudaError_t e2 = cudaSuccess;
  if (!e->done) e2 = cudaEventCreateWithFlags(&e->done, cudaEventDisableTiming);
  else e2 = cudaEventSynchronize(e->done);   // its last launch has finished
  if (e2 == cudaSuccess && e->exec) {
    cudaGraphExecUpdateResultInfo info;
    if (e->fn == fn && e->n_steps == n && e->topo == topo && cudaGraphExecUpdate(e->exec, graph, &info) == cudaSuccess) {
      ++ax.graph_updates;
    } else {
      cudaGetLastError();
      cudaGraphExecDestroy(e->exec);
      e->exec = nullptr;
    }
  }
  if (e2 == cudaSuccess && !e->exec) {
    e2 = cudaGraphInstantiate(&e->exec, graph, 0);
    ++ax.graph_instantiations;
  }
  cudaGraphDestroy(graph);
  if (e2 != cudaSuccess) { e->exec = nullptr; e->key.clear(); set_cuda_error(e2, "CUDA graph build"); return APR_E_CUDA; }
  e->key.assign(reinterpret_cast<const unsigned char*>(&key), reinterpret_cast<const unsigned char*>(&key) + sizeof(StepCtx));
  e->fn = fn; e->n_steps = n; e->topo = topo;
  e->last_use = ++ax.graph_clock;
  *out = e;
  return APR_OK;
}

template <int G, int V, bool FULL>
static int run_steps_t(const StepCtx& c, int mode, cudaStream_t st, const StepDyn* dyn_dev) {
  const int gpb = kThreads / G;
  const int sms = sm_count();
  if (mode == 0) {
    DeviceContext* axp = device_context();
    if (!axp) return APR_E_CUDA;
    DeviceContext& ax = *axp;
    int occ_fast = 0;
    APR_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_fast, fast_kernel<G, V, FULL>, kThreads, 0));
    // resident blocks per SM are shared between the fast kernel and the general-path kernels running next to it
    static const int fast_bps = env_int("APR_FAST_BLOCKS", 0), gen_bps = env_int("APR_GEN_BLOCKS", 0);
    const int fb = fast_bps > 0 ? fast_bps : std::max(1, std::min(2, occ_fast));  // 16 warps/SM already saturate HBM
    const int gb = gen_bps > 0 ? gen_bps : 2;
    const int grid_fast = std::max(1, std::min((c.B + gpb - 1) / gpb, sms * fb));
    const int grid_gen = std::max(1, std::min((c.B + gpb - 1) / gpb, sms * gb));
    if (c.only_stage >= 0) {
      for (int s = c.s_begin; s < c.s_end; ++s) {
        if (c.only_stage == 3) fast_kernel<G, V, FULL><<<grid_fast, kThreads, 0, st>>>(c, s);
        else if (c.only_stage == 4) {
          if constexpr (V == 1) {
            if (c.npair) pair_kernel<G, V><<<std::max(1, std::min((c.B / 2 + gpb - 1) / gpb, sms)), kThreads, 0, st>>>(c, s);
          }
        } else if (c.only_stage > 0 || c.adver) general_stage_kernel<G, V><<<grid_gen, kThreads, 0, st>>>(c, s, c.only_stage);
      }
      APR_LAUNCH_CHECK();
      return APR_OK;
    }
    const bool pairs = V == 1 && c.npair != nullptr;
    // pair blocks hold 128 registers per thread: with 256 threads one block takes half an SM's register file and, next to
    // the two fast-kernel blocks, leaves no room for a general-stage block until it retires (APR_PAIR_THREADS=128: half)
    static const int pair_threads_env = env_int("APR_PAIR_THREADS", 256);
    const int pair_threads = (pair_threads_env == 64 || pair_threads_env == 128) ? pair_threads_env : 256;
    const int gpb_pair = pair_threads / G;
    const int grid_pair = std::max(1, std::min((c.B / 2 + gpb_pair - 1) / gpb_pair, sms * (kThreads / pair_threads)));
    // Which path gets the SM slots first (experiments, APR_GEN_PRIO bit mask; default 0 = the fast and pair kernels run at
    // the highest stream priority and the general stages, on the caller's stream, take what is left):
    //   bit 0  the three dependent general launches run on a highest-priority stream of their own
    //   bit 1  the fast kernel runs at the lowest priority        bit 2  the pair kernel runs at the lowest priority
    // Measured (B=65536, kernel-only fraction): d=64  mask 0: 0.683, 4: 0.684, 1 / 5 / 7: 0.660-0.661;  d=128  mask 0:
    // 0.821, 4: 0.821, 1 / 5 / 7: 0.789-0.791 -- the extra fork/join of a fourth stream costs more than the order buys.
    static const int gen_prio = env_int("APR_GEN_PRIO", 0);
    const bool gen_first = (gen_prio & 1) != 0;
    cudaStream_t fs = (gen_prio & 2) ? ax.fast_lo_stream : ax.fast_stream;
    cudaStream_t ps = (gen_prio & 4) ? ax.pair_lo_stream : ax.pair_stream;
    // the launches of steps [s0, s1) with context cc, forked from / joined to `main`
    auto issue_steps = [&](const StepCtx& cc, cudaStream_t main, int s0, int s1) -> int {
      cudaStream_t gs = gen_first ? ax.gen_stream : main;
      for (int s = s0; s < s1; ++s) {
        APR_CUDA_CHECK(cudaEventRecord(ax.fork, main));
        if (gen_first) {
          APR_CUDA_CHECK(cudaStreamWaitEvent(gs, ax.fork, 0));
          if (cc.adver) general_stage_kernel<G, V><<<grid_gen, kThreads, 0, gs>>>(cc, s, 0);
        }
        APR_CUDA_CHECK(cudaStreamWaitEvent(fs, ax.fork, 0));
        fast_kernel<G, V, FULL><<<grid_fast, kThreads, 0, fs>>>(cc, s);
        APR_CUDA_CHECK(cudaEventRecord(ax.join, fs));
        if (pairs) {
          APR_CUDA_CHECK(cudaStreamWaitEvent(ps, ax.fork, 0));
          if constexpr (V == 1) pair_kernel<G, V><<<grid_pair, pair_threads, 0, ps>>>(cc, s);
          APR_CUDA_CHECK(cudaEventRecord(ax.join2, ps));
        }
        if (!gen_first && cc.adver) general_stage_kernel<G, V><<<grid_gen, kThreads, 0, gs>>>(cc, s, 0);
        general_stage_kernel<G, V><<<grid_gen, kThreads, 0, gs>>>(cc, s, 1);
        general_stage_kernel<G, V><<<grid_gen, kThreads, 0, gs>>>(cc, s, 2);
        if (gen_first) {
          APR_CUDA_CHECK(cudaEventRecord(ax.join3, gs));
          APR_CUDA_CHECK(cudaStreamWaitEvent(main, ax.join3, 0));
        }
        APR_CUDA_CHECK(cudaStreamWaitEvent(main, ax.join, 0));
        if (pairs) APR_CUDA_CHECK(cudaStreamWaitEvent(main, ax.join2, 0));
      }
      APR_LAUNCH_CHECK();
      return APR_OK;
    };
    // Graph replay pays off when a step is long enough for the saved dependency latency to matter and the batch is large
    // enough for the three-stream layout (APR_GRAPH=0 switches it off, APR_GRAPH_MIN_BATCH moves the threshold).
    static const int use_graph = env_int("APR_GRAPH", 1), graph_min_batch = env_int("APR_GRAPH_MIN_BATCH", 1024);
    if (use_graph && c.B >= graph_min_batch && dyn_dev) {
      StepCtx key = c;                    // static part: everything except the steps and the stats pointer of this call
      key.s_begin = key.s_end = 0;
      key.stats = nullptr;
      key.dyn = dyn_dev;
      const void* fn = reinterpret_cast<const void*>(&fast_kernel<G, V, FULL>);
      const int topo = (c.adver ? 1 : 0) | (pairs ? 2 : 0);
      auto issue_group = [&](cudaStream_t main, int n) -> int {
        const int rc = issue_steps(key, main, 0, n);
        if (rc) return rc;
        dyn_advance_kernel<<<1, 1, 0, main>>>(const_cast<StepDyn*>(dyn_dev), n);
        APR_LAUNCH_CHECK();
        return APR_OK;
      };
      constexpr size_t kNG = sizeof(kGraphGroups) / sizeof(int);
      GraphEntry* by_size[kNG];
      for (size_t k = 0; k < kNG; ++k) {
        by_size[k] = graph_find(ax, key, fn, kGraphGroups[k]);
        if (!by_size[k]) {
          const int rc = graph_build(ax, key, fn, kGraphGroups[k], topo, issue_group, &by_size[k]);
          if (rc) return rc;
        }
      }
      // the cursor carries over from one sub-chunk of a call to the next (every graph advances it by its own step
      // count): it is set once, in front of the call's first sub-chunk
      if (c.s_begin == 0) {
        dyn_set_kernel<<<1, 1, 0, st>>>(const_cast<StepDyn*>(dyn_dev), 0, c.stats);
        APR_LAUNCH_CHECK();
      }
      int left = c.s_end - c.s_begin;
      for (size_t k = 0; k < kNG; ++k) {
        while (left >= kGraphGroups[k]) {
          GraphEntry* e = by_size[k];
          APR_CUDA_CHECK(cudaGraphLaunch(e->exec, st));
          APR_CUDA_CHECK(cudaEventRecord(e->done, st));
          e->last_use = ++ax.graph_clock;
          ++ax.graph_launches;
          left -= kGraphGroups[k];
        }
      }
      return APR_OK;
    }
    return issue_steps(c, st, c.s_begin, c.s_end);
  }
  if (mode == 2) {
    // small batches (tables in L2, steps of a few microseconds): ONE cluster of up to 16 CTAs runs every step of the
    // call; a cluster barrier costs a fraction of a microsecond where a kernel boundary or a grid barrier costs several
    static const int cl_max = env_int("APR_CLUSTER_BLOCKS", 16);
    const int grid = std::max(1, std::min(std::min(cl_max, 16), (c.B + gpb - 1) / gpb));
    StepCtx cc = c;
    cc.cluster_sync = 1;
    if (grid > 8)
      APR_CUDA_CHECK(cudaFuncSetAttribute(step_persistent_kernel<G, V, FULL>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = grid;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    APR_CUDA_CHECK(cudaLaunchKernelEx(&cfg, step_persistent_kernel<G, V, FULL>, cc));
    return APR_OK;
  }
  int occ = 0;
  APR_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, step_persistent_kernel<G, V, FULL>, kThreads, 0));
  if (occ < 1) return APR_E_CUDA;
  int grid = sms * occ;
  grid = std::max(1, std::min(grid, (c.B + gpb - 1) / gpb));
  StepCtx cc = c;
  void* args[] = {&cc};
  APR_CUDA_CHECK(
      cudaLaunchCooperativeKernel((void*)step_persistent_kernel<G, V, FULL>, dim3(grid), dim3(kThreads), args, 0, st));
  return APR_OK;
}

template <int G, int V>
static int run_steps(const StepCtx& c, int mode, cudaStream_t st, const StepDyn* dyn_dev) {
  if (c.d == G * V * 4) return run_steps_t<G, V, true>(c, mode, st, dyn_dev);
  return run_steps_t<G, V, false>(c, mode, st, dyn_dev);
}

// dyn_dev: device StepDyn of the workspace (graph replay allowed) or nullptr (plain launches only)
static int dispatch_steps(const StepCtx& c, int mode, cudaStream_t st, const StepDyn* dyn_dev) {
  const int q = c.d / 4;
  if (q <= 4) return run_steps<4, 1>(c, mode, st, dyn_dev);
  if (q <= 8) return run_steps<8, 1>(c, mode, st, dyn_dev);
  if (q <= 16) return run_steps<16, 1>(c, mode, st, dyn_dev);
  if (q <= 32) return run_steps<32, 1>(c, mode, st, dyn_dev);
  if (q <= 64) return run_steps<32, 2>(c, mode, st, dyn_dev);
  if (q <= 96) return run_steps<32, 3>(c, mode, st, dyn_dev);
  return run_steps<32, 4>(c, mode, st, dyn_dev);
}

// ------------------------------------------------------------------------------------------------
// loss / accuracy (utils.py:159-175): one CTA per (step, 256-triple tile)
// ------------------------------------------------------------------------------------------------
template <int G, int V>
__global__ void __launch_bounds__(kThreads) loss_acc_kernel(const float* __restrict__ P, const float* __restrict__ Q, int d,
                                                            const int32_t* __restrict__ u, const int32_t* __restrict__ i,
                                                            const int32_t* __restrict__ j, int B, int tiles_per_step,
                                                            double* out) {
  constexpr int kTile = 256;
  const int s = blockIdx.x / tiles_per_step;
  const int tile = blockIdx.x - s * tiles_per_step;
  const int lane = threadIdx.x % G, g = threadIdx.x / G;
  const unsigned mask = group_mask<G>();
  const int b_end = min(B, (tile + 1) * kTile);
  float loss = 0.f;
  int correct = 0;
  for (int b = tile * kTile + g; b < b_end; b += kThreads / G) {
    const int64_t t = int64_t(s) * B + b;
    Row<G, V> p, q, n;
    row_load<G, V>(p, P + int64_t(u[t]) * d, lane, d);
    row_load<G, V>(q, Q + int64_t(i[t]) * d, lane, d);
    row_load<G, V>(n, Q + int64_t(j[t]) * d, lane, d);
    const float x = row_dot<G, V>(p, q, mask) - row_dot<G, V>(p, n, mask);
    const float r = fminf(fmaxf(x, -80.0f), 1e8f);
    loss += softplus_neg(r);
    correct += x > 0.f;
  }
  __shared__ double sh_loss;
  __shared__ int sh_cor;
  if (threadIdx.x == 0) { sh_loss = 0.0; sh_cor = 0; }
  __syncthreads();
  if (lane == 0) { atomicAdd(&sh_loss, double(loss)); atomicAdd(&sh_cor, correct); }
  __syncthreads();
  if (threadIdx.x == 0) { atomicAdd(&out[2 * s], sh_loss); atomicAdd(&out[2 * s + 1], double(sh_cor)); }
}

template <int G, int V>
static int run_loss_acc(const float* P, const float* Q, int d, const int32_t* u, const int32_t* i, const int32_t* j, int S,
                        int B, double* out, cudaStream_t st) {
  const int tiles = (B + 255) / 256;
  loss_acc_kernel<G, V><<<S * tiles, kThreads, 0, st>>>(P, Q, d, u, i, j, B, tiles, out);
  APR_LAUNCH_CHECK();
  return APR_OK;
}

// index preparation of the steps [s0, s0+ns) (one L2-sized sub-chunk) on stream st
// `pairs`: also detect pair work units (d <= 128)
static int prepare_sub(const int32_t* u, const int32_t* i, const int32_t* j, const TrainLayout& L, int s0, int ns,
                       int64_t rows_p, int64_t rows_q, void* ws, cudaStream_t st, bool pairs) {
  const int B = L.B;
  int32_t* tkey_u = at<int32_t>(ws, L.off_tkey_u);
  int32_t* tval_u = at<int32_t>(ws, L.off_tval_u);
  int32_t* tkey_i = at<int32_t>(ws, L.off_tkey_i);
  int32_t* tval_i = at<int32_t>(ws, L.off_tval_i);
  const int threads = 256;
  // resident preparation blocks per SM: the preparation must finish a sub-chunk faster than the steps consume one, or the
  // step pipeline stalls (B=65536, d=128: 2 -> 638, 4 -> 723, 8 -> 758, 16 -> 762, 32 -> 768 M triples/s)
  static const int prep_bps = env_int("APR_PREP_BLOCKS", 32);
  const int64_t cap = int64_t(sm_count()) * std::max(1, prep_bps);
  // table keys = -1, table values = 0
  APR_CUDA_CHECK(cudaMemsetAsync(tkey_u, 0xFF, size_t(int64_t(ns) * L.Tu * 4), st));
  APR_CUDA_CHECK(cudaMemsetAsync(tval_u, 0, size_t(int64_t(ns) * L.Tu * 4), st));
  APR_CUDA_CHECK(cudaMemsetAsync(tkey_i, 0xFF, size_t(int64_t(ns) * L.Ti * 4), st));
  APR_CUDA_CHECK(cudaMemsetAsync(tval_i, 0, size_t(int64_t(ns) * L.Ti * 4), st));
  const int64_t nt = int64_t(ns) * B;
  const int grid_a = int(std::max<int64_t>(1, std::min<int64_t>((nt + threads - 1) / threads, cap)));
  prep_insert_kernel<<<grid_a, threads, 0, st>>>(u, i, j, s0, ns, B, rows_p, rows_q, L.Tu, L.Ti, tkey_u, tval_u, tkey_i,
                                                 tval_i, at<int32_t>(ws, L.off_entry), at<int32_t>(ws, L.off_hdr));
  const int64_t ne = int64_t(ns) * (L.Tu + L.Ti);
  const int grid_b = int(std::max<int64_t>(1, std::min<int64_t>((ne + threads - 1) / threads, cap)));
  static const int compact_block = env_int("APR_COMPACT_BLOCK", 1);
  if (compact_block && L.Tu % 256 == 0 && L.Ti % 256 == 0)
    prep_compact_block_kernel<<<grid_b, 256, 0, st>>>(
        s0, ns, B, L.Tu, L.Ti, tkey_u, tval_u, tkey_i, tval_i, at<int32_t>(ws, L.off_ucnt), at<int32_t>(ws, L.off_icnt),
        at<int32_t>(ws, L.off_iall), at<int32_t>(ws, L.off_tcursor), at<int32_t>(ws, L.off_seg_user),
        at<int32_t>(ws, L.off_seg_off), at<int32_t>(ws, L.off_seg_cnt), at<int32_t>(ws, L.off_ucursor),
        at<int32_t>(ws, L.off_iu_item));
  else
    prep_compact_kernel<<<grid_b, threads, 0, st>>>(
        s0, ns, B, L.Tu, L.Ti, tkey_u, tval_u, tkey_i, tval_i, at<int32_t>(ws, L.off_ucnt), at<int32_t>(ws, L.off_icnt),
        at<int32_t>(ws, L.off_iall), at<int32_t>(ws, L.off_tcursor), at<int32_t>(ws, L.off_seg_user),
        at<int32_t>(ws, L.off_seg_off), at<int32_t>(ws, L.off_seg_cnt), at<int32_t>(ws, L.off_ucursor),
        at<int32_t>(ws, L.off_iu_item));
  prep_scatter_kernel<<<grid_a, threads, 0, st>>>(i, j, s0, ns, B, L.Tu, L.Ti, tval_u, tval_i, at<int32_t>(ws, L.off_entry),
                                                  at<int32_t>(ws, L.off_seg_off), at<int32_t>(ws, L.off_ucursor),
                                                  at<int4>(ws, L.off_rec), at<int32_t>(ws, L.off_seg_slow),
                                                  pairs ? at<int32_t>(ws, L.off_occ_cnt) : nullptr,
                                                  at<int2>(ws, L.off_occ_seg));
  if (pairs)
    prep_pair_kernel<<<grid_a, threads, 0, st>>>(s0, ns, B, at<int32_t>(ws, L.off_icnt), at<int32_t>(ws, L.off_occ_cnt),
                                                 at<int2>(ws, L.off_occ_seg), at<int32_t>(ws, L.off_seg_user),
                                                 at<int32_t>(ws, L.off_seg_off), at<int32_t>(ws, L.off_seg_cnt),
                                                 at<int32_t>(ws, L.off_seg_slow), at<int4>(ws, L.off_rec),
                                                 at<int32_t>(ws, L.off_iu_item), at<int32_t>(ws, L.off_npair),
                                                 at<int4>(ws, L.off_pairs));
  if (compact_block && B % 256 == 0)
    prep_pack_block_kernel<<<grid_a, 256, 0, st>>>(s0, ns, B, at<int32_t>(ws, L.off_ucnt), at<int32_t>(ws, L.off_seg_user),
                                                   at<int32_t>(ws, L.off_seg_off), at<int32_t>(ws, L.off_seg_cnt),
                                                   at<int32_t>(ws, L.off_seg_slow), at<int4>(ws, L.off_rec),
                                                   at<int4>(ws, L.off_seg_hdr), at<int32_t>(ws, L.off_nslow),
                                                   at<int32_t>(ws, L.off_nfast));
  else
    prep_pack_kernel<<<grid_a, threads, 0, st>>>(s0, ns, B, at<int32_t>(ws, L.off_ucnt), at<int32_t>(ws, L.off_seg_user),
                                                 at<int32_t>(ws, L.off_seg_off), at<int32_t>(ws, L.off_seg_cnt),
                                                 at<int32_t>(ws, L.off_seg_slow), at<int4>(ws, L.off_rec),
                                                 at<int4>(ws, L.off_seg_hdr), at<int32_t>(ws, L.off_nslow),
                                                 at<int32_t>(ws, L.off_nfast));
  APR_LAUNCH_CHECK();
  return APR_OK;
}

// per-step counters and slow flags of the steps [s0, s0+ns) = 0 (other steps' arrays may hold another rank's broadcast, or
// simply are not part of this call: the workspace may have room for many more steps than a call uses)
static int prepare_clear_range(const TrainLayout& L, void* ws, int s0, int ns, cudaStream_t st) {
  const int64_t offs[7] = {L.off_ucnt, L.off_icnt, L.off_iall, L.off_nslow, L.off_nfast, L.off_tcursor, L.off_npair};
  for (int k = 0; k < 7; ++k) APR_CUDA_CHECK(cudaMemsetAsync(at<char>(ws, offs[k]) + int64_t(s0) * 4, 0, size_t(ns) * 4, st));
  APR_CUDA_CHECK(cudaMemsetAsync(at<char>(ws, L.off_seg_slow) + int64_t(s0) * L.B * 4, 0, size_t(ns) * L.B * 4, st));
  APR_CUDA_CHECK(cudaMemsetAsync(at<char>(ws, L.off_occ_cnt) + int64_t(s0) * L.B * 4, 0, size_t(ns) * L.B * 4, st));
  return APR_OK;
}

static int prepare_impl(const int32_t* u, const int32_t* i, const int32_t* j, int S, int B, int d, int64_t rows_p,
                        int64_t rows_q, void* ws, int64_t ws_bytes, cudaStream_t st) {
  TrainLayout L;
  if (!layout_for_ws(ws_bytes, B, d, S, L)) return APR_E_WORKSPACE;
  int rc = prepare_clear_range(L, ws, 0, S, st);
  for (int s0 = 0; s0 < S && !rc; s0 += L.Sc)
    rc = prepare_sub(u, i, j, L, s0, std::min(L.Sc, S - s0), rows_p, rows_q, ws, st, pairs_enabled(d));
  return rc;
}

}  // namespace apr

using namespace apr;

extern "C" {

int64_t apr_train_workspace_bytes(int32_t n_steps, int32_t batch, int32_t d) {
  if (n_steps < 1 || batch < 1 || !valid_dim(d)) return -1;
  return make_layout(n_steps, batch, d).total;
}

int apr_train_workspace_init(void* ws, int64_t ws_bytes, apr_stream_t stream) {
  if (!ws || ws_bytes <= 0) return APR_E_ARG;
  APR_CUDA_CHECK(cudaMemsetAsync(ws, 0, size_t(ws_bytes), static_cast<cudaStream_t>(stream)));
  return APR_OK;
}

int apr_train_prepare(const int32_t* u, const int32_t* i, const int32_t* j, int32_t S, int32_t B, int32_t d, int64_t rows_p,
                      int64_t rows_q, void* ws, int64_t ws_bytes, apr_stream_t stream) {
  if (!u || !i || !j || !ws || S < 1 || B < 1 || rows_p < 1 || rows_q < 1 || !valid_dim(d)) return APR_E_ARG;
  if (!aligned16(ws)) return APR_E_ALIGN;
  return prepare_impl(u, i, j, S, B, d, rows_p, rows_q, ws, ws_bytes, static_cast<cudaStream_t>(stream));
}

static int check_train_args(const float* P, const float* Q, const float* accP, const float* accQ, int64_t rows_p,
                            int64_t rows_q, int d, const int32_t* u, const int32_t* i, const int32_t* j, int S, int B,
                            const void* ws, int adver) {
  if (!P || !Q || !accP || !accQ || !u || !i || !j || !ws) return APR_E_ARG;
  if (adver < 0 || adver > 3) return APR_E_ARG;
  if (adver == 2) return APR_E_UNSUPPORTED;   // --adv random goes through apr_train_steps_random
  if (S < 1 || B < 1 || rows_p < 1 || rows_q < 1 || !valid_dim(d)) return APR_E_ARG;
  if (!aligned16(P) || !aligned16(Q) || !aligned16(accP) || !aligned16(accQ) || !aligned16(ws)) return APR_E_ALIGN;
  return APR_OK;
}

static int run_range(float* P, float* Q, float* accP, float* accQ, int32_t d, int32_t S, int32_t B, float lr, float reg,
                     float reg_adv, float eps, int32_t adver, int32_t mode, void* ws, const TrainLayout& L, float* stats,
                     int s_begin, int s_end, cudaStream_t st, uint32_t noise_seed = 0, uint32_t noise_step0 = 0) {
  StepCtx c;
  memset(&c, 0, sizeof(c));
  c.Pb[0] = P; c.Qb[0] = Q; c.aPb[0] = accP; c.aQb[0] = accQ;
  c.GQb[0] = at<float>(ws, L.off_GQ); c.HQb[0] = at<float>(ws, L.off_HQ);
  c.nranks = 1; c.rank = 0; c.rshift = 0; c.only_stage = -1; c.cluster_sync = 0;
  c.d = d; c.B = B; c.S = L.S;
  c.lr = lr;
  // k = 2 reg (1 + [adver]) / (B d): the mean-regulariser is added once, or twice when adver (APR.py:153-154,163-165)
  c.kreg = float(2.0 * double(reg) * (adver ? 2.0 : 1.0) / (double(B) * double(d)));
  // adver 3 = the reference's dns > 1 branch on an adversarial graph (utils.py:121-139): the optimizer minimises
  // L + reg_adv L_adv + 2 reg-terms with Delta == 0 (update_P / update_Q never run there), i.e. a plain step whose data
  // term is scaled by 1 + reg_adv and whose regulariser is counted twice
  c.reg_adv = reg_adv; c.eps = eps; c.adver = adver == 1 ? 1 : 0;
  c.plain_scale = adver == 3 ? 1.0f + reg_adv : 1.0f;
  c.ucnt = at<int32_t>(ws, L.off_ucnt); c.icnt = at<int32_t>(ws, L.off_icnt);
  c.nslow = at<int32_t>(ws, L.off_nslow);
  c.npair = pairs_enabled(d) ? at<int32_t>(ws, L.off_npair) : nullptr;
  c.pairs = at<int4>(ws, L.off_pairs);
  c.seg_hdr = at<int4>(ws, L.off_seg_hdr);
  c.rec = at<int4>(ws, L.off_rec);
  c.iu_item = at<int32_t>(ws, L.off_iu_item);
  c.GP = at<float>(ws, L.off_GP);
  c.cbuf = at<float>(ws, L.off_cbuf);
  c.QN = at<float>(ws, L.off_QN);
  c.stats = stats;
  c.flags = env_int("APR_STEP_FLAGS", 0);
  c.s_begin = s_begin; c.s_end = s_end;
  if (adver == 2) {
    c.noise_seed = noise_seed; c.noise_step0 = noise_step0;
    return dispatch_random_steps(c, st);
  }
  return dispatch_steps(c, mode, st, reinterpret_cast<const StepDyn*>(at<char>(ws, L.off_hdr) + kDynOffset));
}

int apr_train_run(float* P, float* Q, float* accP, float* accQ, int64_t rows_p, int64_t rows_q, int32_t d,
                  const int32_t* u, const int32_t* i, const int32_t* j, int32_t S, int32_t B, float lr, float reg,
                  float reg_adv, float eps, int32_t adver, int32_t mode, void* ws, int64_t ws_bytes, float* stats,
                  apr_stream_t stream) {
  int rc = check_train_args(P, Q, accP, accQ, rows_p, rows_q, d, u, i, j, S, B, ws, adver);
  if (rc) return rc;
  if (mode < 0 || mode > 2) return APR_E_ARG;
  TrainLayout L;
  if (!layout_for_ws(ws_bytes, B, d, S, L)) return APR_E_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DeviceContext* ctx = device_context();
  if (!ctx) return APR_E_CUDA;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  if (stats) APR_CUDA_CHECK(cudaMemsetAsync(stats, 0, size_t(S) * 2 * sizeof(float), st));
  const int sub = mode == 0 ? L.Sc : S;
  for (int s0 = 0; s0 < S; s0 += sub) {
    rc = run_range(P, Q, accP, accQ, d, S, B, lr, reg, reg_adv, eps, adver, mode, ws, L, stats, s0, std::min(S, s0 + sub), st);
    if (rc) return rc;
  }
  return APR_OK;
}

// prepare + run, software-pipelined: the index preparation of sub-chunk c+1 runs on its own stream while the step
// kernels of sub-chunk c run on the caller's stream (prepare touches only its hash tables and the per-step arrays of its
// own steps; the step kernels touch only the embedding tables, the gradient workspace and their own steps' arrays).
int apr_train_steps(float* P, float* Q, float* accP, float* accQ, int64_t rows_p, int64_t rows_q, int32_t d,
                    const int32_t* u, const int32_t* i, const int32_t* j, int32_t S, int32_t B, float lr, float reg,
                    float reg_adv, float eps, int32_t adver, int32_t mode, void* ws, int64_t ws_bytes, float* stats,
                    apr_stream_t stream) {
  int rc = check_train_args(P, Q, accP, accQ, rows_p, rows_q, d, u, i, j, S, B, ws, adver);
  if (rc) return rc;
  if (mode < 0 || mode > 2) return APR_E_ARG;
  TrainLayout L;
  if (!layout_for_ws(ws_bytes, B, d, S, L)) return APR_E_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DeviceContext* ctx = device_context();
  if (!ctx) return APR_E_CUDA;
  DeviceContext& ax = *ctx;
  std::lock_guard<std::recursive_mutex> lk(ax.mu);
  if (stats) APR_CUDA_CHECK(cudaMemsetAsync(stats, 0, size_t(S) * 2 * sizeof(float), st));
  // everything already enqueued on the caller's stream (producers of u,i,j; earlier steps reading the arrays) first
  APR_CUDA_CHECK(cudaEventRecord(ax.entry, st));
  APR_CUDA_CHECK(cudaStreamWaitEvent(ax.prep_stream, ax.entry, 0));
  rc = prepare_clear_range(L, ws, 0, S, ax.prep_stream);
  if (rc) return rc;
  // sub-chunks of 2, 4, 8, ... up to Sc steps: only the index preparation of the FIRST sub-chunk is exposed in front of
  // the step kernels (it cannot start before the previous call's steps have released the per-step arrays), so it is
  // kept short; every later one runs under the steps of its predecessor
  std::vector<int> sizes;
  for (int left = S, n = std::min(2, L.Sc); left > 0; n = std::min(L.Sc, 2 * n)) {
    sizes.push_back(std::min(left, n));
    left -= sizes.back();
  }
  for (size_t k = 0, s0 = 0; k < sizes.size(); s0 += sizes[k], ++k) {
    rc = prepare_sub(u, i, j, L, int(s0), sizes[k], rows_p, rows_q, ws, ax.prep_stream, pairs_enabled(d));
    if (rc) return rc;
    cudaEvent_t e = ax.prep_event(k);
    if (!e) return APR_E_CUDA;
    APR_CUDA_CHECK(cudaEventRecord(e, ax.prep_stream));
  }
  for (size_t k = 0, s0 = 0; k < sizes.size(); s0 += sizes[k], ++k) {
    APR_CUDA_CHECK(cudaStreamWaitEvent(st, ax.prep_event(k), 0));
    rc = run_range(P, Q, accP, accQ, d, S, B, lr, reg, reg_adv, eps, adver, mode, ws, L, stats, int(s0), int(s0) + sizes[k], st);
    if (rc) return rc;
  }
  return APR_OK;
}

// --adv random (APR.py:170-177): see random_step_kernel.  Steps s = 0..S-1 of this call use the noise of global step
// noise_step0 + s; the caller advances noise_step0 by S from call to call.
int apr_train_steps_random(float* P, float* Q, float* accP, float* accQ, int64_t rows_p, int64_t rows_q, int32_t d,
                           const int32_t* u, const int32_t* i, const int32_t* j, int32_t S, int32_t B, float lr, float reg,
                           float reg_adv, float eps, uint32_t noise_seed, uint32_t noise_step0, void* ws, int64_t ws_bytes,
                           float* stats, apr_stream_t stream) {
  int rc = check_train_args(P, Q, accP, accQ, rows_p, rows_q, d, u, i, j, S, B, ws, 1);
  if (rc) return rc;
  TrainLayout L;
  if (!layout_for_ws(ws_bytes, B, d, S, L)) return APR_E_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DeviceContext* ctx = device_context();
  if (!ctx) return APR_E_CUDA;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  if (stats) APR_CUDA_CHECK(cudaMemsetAsync(stats, 0, size_t(S) * 2 * sizeof(float), st));
  rc = prepare_impl(u, i, j, S, B, d, rows_p, rows_q, ws, ws_bytes, st);
  if (rc) return rc;
  return run_range(P, Q, accP, accQ, d, S, B, lr, reg, reg_adv, eps, 2, 0, ws, L, stats, 0, S, st, noise_seed, noise_step0);
}

int apr_train_layout(int32_t S, int32_t B, int32_t d, int64_t* out) {  // 13 entries
  if (!out || S < 1 || B < 1 || !valid_dim(d)) return APR_E_ARG;
  const TrainLayout L = make_layout(S, B, d);
  out[0] = L.total; out[1] = L.Sc; out[2] = L.off_ucnt; out[3] = L.off_icnt; out[4] = L.off_iall; out[5] = L.off_nslow;
  out[6] = L.off_seg_hdr; out[7] = L.off_rec; out[8] = L.off_iu_item; out[9] = L.off_hdr;
  out[10] = L.off_npair; out[11] = L.off_nfast; out[12] = L.off_pairs;
  return APR_OK;
}

int apr_train_prepare_range(const int32_t* u, const int32_t* i, const int32_t* j, int32_t S, int32_t B, int32_t d,
                            int64_t rows_p, int64_t rows_q, void* ws, int64_t ws_bytes, int32_t s0, int32_t ns,
                            int32_t clear_counters, apr_stream_t stream) {
  if (!u || !i || !j || !ws || S < 1 || B < 1 || rows_p < 1 || rows_q < 1 || !valid_dim(d)) return APR_E_ARG;
  if (s0 < 0 || ns < 1 || s0 + ns > S) return APR_E_ARG;
  if (!aligned16(ws)) return APR_E_ALIGN;
  TrainLayout L;
  if (!layout_for_ws(ws_bytes, B, d, S, L)) return APR_E_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc = clear_counters ? prepare_clear_range(L, ws, s0, ns, st) : APR_OK;
  for (int a = s0; a < s0 + ns && !rc; a += L.Sc) rc = prepare_sub(u, i, j, L, a, std::min(L.Sc, s0 + ns - a), rows_p, rows_q, ws, st, pairs_enabled(d));
  return rc;
}

static int stage_sharded_impl(float* const* Pb, float* const* Qb, float* const* accPb, float* const* accQb,
                              float* const* GQb, float* const* HQb, int32_t nranks, int32_t rank, int32_t d, int32_t S,
                              int32_t B, float lr, float reg, float reg_adv, float eps, int32_t adver, void* ws,
                              int64_t ws_bytes, float* stats, int32_t step, int32_t stage, const int* abort_flag,
                              apr_stream_t stream) {
  if (!Pb || !Qb || !accPb || !accQb || !GQb || !HQb || !ws) return APR_E_ARG;
  if (adver < 0 || adver > 1) return APR_E_UNSUPPORTED;   // row-sharded tables: BPR and gradient-based APR only
  if (nranks < 1 || nranks > kMaxRanks || (nranks & (nranks - 1)) || rank < 0 || rank >= nranks) return APR_E_ARG;
  if (S < 1 || B < 1 || !valid_dim(d) || step < 0 || step >= S || stage < 0 || stage > 4) return APR_E_ARG;
  TrainLayout L;
  if (!layout_for_ws(ws_bytes, B, d, S, L)) return APR_E_WORKSPACE;
  DeviceContext* ctx = device_context();
  if (!ctx) return APR_E_CUDA;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  StepCtx c;
  memset(&c, 0, sizeof(c));
  for (int r = 0; r < nranks; ++r) {
    if (!Pb[r] || !Qb[r] || !accPb[r] || !accQb[r] || !GQb[r] || !HQb[r]) return APR_E_ARG;
    if (!aligned16(Pb[r]) || !aligned16(Qb[r]) || !aligned16(accPb[r]) || !aligned16(accQb[r]) || !aligned16(GQb[r]) ||
        !aligned16(HQb[r]))
      return APR_E_ALIGN;
    c.Pb[r] = Pb[r]; c.Qb[r] = Qb[r]; c.aPb[r] = accPb[r]; c.aQb[r] = accQb[r]; c.GQb[r] = GQb[r]; c.HQb[r] = HQb[r];
  }
  c.nranks = nranks; c.rank = rank;
  c.rshift = 0;
  while ((1 << c.rshift) < nranks) ++c.rshift;
  c.d = d; c.B = B; c.S = L.S; c.lr = lr;
  c.kreg = float(2.0 * double(reg) * (adver ? 2.0 : 1.0) / (double(B) * double(d)));
  c.reg_adv = reg_adv; c.eps = eps; c.adver = adver ? 1 : 0;
  c.plain_scale = 1.0f;
  c.abort = abort_flag;
  c.ucnt = at<int32_t>(ws, L.off_ucnt); c.icnt = at<int32_t>(ws, L.off_icnt); c.nslow = at<int32_t>(ws, L.off_nslow);
  c.npair = pairs_enabled(d) ? at<int32_t>(ws, L.off_npair) : nullptr; c.pairs = at<int4>(ws, L.off_pairs);
  c.seg_hdr = at<int4>(ws, L.off_seg_hdr); c.rec = at<int4>(ws, L.off_rec); c.iu_item = at<int32_t>(ws, L.off_iu_item);
  c.GP = at<float>(ws, L.off_GP); c.cbuf = at<float>(ws, L.off_cbuf); c.QN = at<float>(ws, L.off_QN);
  c.stats = stats;
  c.flags = env_int("APR_STEP_FLAGS", 0);
  c.s_begin = step; c.s_end = step + 1; c.only_stage = stage; c.cluster_sync = 0;
  return dispatch_steps(c, 0, static_cast<cudaStream_t>(stream), nullptr);
}

int apr_train_stage_sharded(float* const* Pb, float* const* Qb, float* const* accPb, float* const* accQb,
                            float* const* GQb, float* const* HQb, int32_t nranks, int32_t rank, int32_t d, int32_t S,
                            int32_t B, float lr, float reg, float reg_adv, float eps, int32_t adver, void* ws,
                            int64_t ws_bytes, float* stats, int32_t step, int32_t stage, apr_stream_t stream) {
  return stage_sharded_impl(Pb, Qb, accPb, accQb, GQb, HQb, nranks, rank, d, S, B, lr, reg, reg_adv, eps, adver, ws, ws_bytes,
                            stats, step, stage, nullptr, stream);
}

// Cross-rank barrier on peer-mapped signal words: rank r stores `epoch` into slot [r] of every peer's signal array and
// waits until all of its own slots reach `epoch`.  Stream ordered (one tiny kernel); spins are bounded (~2 s) and set
// *err instead of hanging.  Everything the earlier kernels of this stream wrote (also to peer memory) is complete at the
// kernel boundary; the system-scope fences order it with the signal.  The pointer array travels as a kernel parameter.
struct SigArray { int* p[kMaxRanks]; };
__global__ void xbarrier_launch(SigArray sig, int nranks, int rank, int epoch, int* err) {
  const int t = threadIdx.x;
  if (t >= nranks) return;
  __threadfence_system();
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(sig.p[t] + rank), "r"(epoch) : "memory");
  const int* src = sig.p[rank] + t;
  const long long t0 = clock64();
  int v;
  do {
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(src) : "memory");
    if (v >= epoch) break;
    // a peer that is ~2 s late (or an earlier time-out of this rank): set *err -- every later stage kernel of this call
    // then returns at once (StepCtx::abort) instead of reading tables the peers are still writing, and the host raises
    if (*reinterpret_cast<volatile int*>(err) || clock64() - t0 > 4000000000LL) { atomicExch(err, 1); break; }
    __nanosleep(64);
  } while (true);
  __threadfence_system();
}

// Steps [first_step, first_step+count) on row-sharded tables, everything launched from here: per step the fast kernel on
// the second stream, the general stages on the caller's stream with a cross-rank barrier after the plain stage, after
// the adversarial stage and at the end of the step.  sig = nranks peer pointers to >= nranks ints each (zeroed once);
// err = device int32 of this rank.  Every rank must make the same sequence of calls.
int apr_train_steps_sharded(float* const* Pb, float* const* Qb, float* const* accPb, float* const* accQb,
                            float* const* GQb, float* const* HQb, int32_t* const* sig, int32_t nranks, int32_t rank,
                            int32_t d, int32_t S, int32_t B, float lr, float reg, float reg_adv, float eps, int32_t adver,
                            void* ws, int64_t ws_bytes, float* stats, int32_t first_step, int32_t count, int32_t* err,
                            apr_stream_t stream) {
  if (!sig || !err || count < 1 || first_step < 0 || first_step + count > S) return APR_E_ARG;
  DeviceContext* ctx = device_context();
  if (!ctx) return APR_E_CUDA;
  DeviceContext& ax = *ctx;
  std::lock_guard<std::recursive_mutex> lk(ax.mu);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SigArray sa;
  for (int r = 0; r < kMaxRanks; ++r) sa.p[r] = r < nranks ? sig[r] : nullptr;
  auto barrier = [&]() -> int {
    if (nranks == 1) return APR_OK;
    xbarrier_launch<<<1, 32, 0, st>>>(sa, nranks, rank, ++ax.xbarrier_epoch, err);
    APR_LAUNCH_CHECK();
    return APR_OK;
  };
  // APR_SHARD_TIMING=1: per-stage device times (events) of this call, printed by rank 0 -- diagnosis only
  static const int timing = env_int("APR_SHARD_TIMING", 0);
  cudaEvent_t ev[8];
  float acc_ms[7] = {0, 0, 0, 0, 0, 0, 0};
  if (timing) for (auto& e : ev) cudaEventCreate(&e);
  for (int s = first_step; s < first_step + count; ++s) {
    int rc;
    // where the fast kernel (second stream) starts relative to the general stages: 0 = with stage 0, 1 = after stage 0
    // and its barrier (the general path's dependent peer loads then do not queue behind the fast kernel's NVLink
    // traffic during stage 0), 2 = after the whole general path (no overlap)
    static const int order = env_int("APR_SHARD_ORDER", 0);
    auto launch_fast = [&]() -> int {
      APR_CUDA_CHECK(cudaEventRecord(ax.fork, st));
      APR_CUDA_CHECK(cudaStreamWaitEvent(ax.fast_stream, ax.fork, 0));
      if (timing) cudaEventRecord(ev[6], ax.fast_stream);
      const int r2 = stage_sharded_impl(Pb, Qb, accPb, accQb, GQb, HQb, nranks, rank, d, S, B, lr, reg, reg_adv, eps,
                                        adver, ws, ws_bytes, stats, s, 3, err, ax.fast_stream);
      if (r2) return r2;
      APR_CUDA_CHECK(cudaEventRecord(ax.join, ax.fast_stream));
      if (timing) cudaEventRecord(ev[7], ax.fast_stream);
      if (pairs_enabled(d)) {   // pair work units: third stream, no ordering against the other kernels of the step
        APR_CUDA_CHECK(cudaStreamWaitEvent(ax.pair_stream, ax.fork, 0));
        const int r3 = stage_sharded_impl(Pb, Qb, accPb, accQb, GQb, HQb, nranks, rank, d, S, B, lr, reg, reg_adv, eps,
                                          adver, ws, ws_bytes, stats, s, 4, err, ax.pair_stream);
        if (r3) return r3;
        APR_CUDA_CHECK(cudaEventRecord(ax.join2, ax.pair_stream));
      }
      return APR_OK;
    };
    if (order == 0 || !adver) { if ((rc = launch_fast())) return rc; }
    if (timing) cudaEventRecord(ev[0], st);
    if (adver) {
      rc = stage_sharded_impl(Pb, Qb, accPb, accQb, GQb, HQb, nranks, rank, d, S, B, lr, reg, reg_adv, eps, adver, ws,
                              ws_bytes, stats, s, 0, err, st);
      if (rc) return rc;
      if (timing) cudaEventRecord(ev[1], st);
      if ((rc = barrier())) return rc;
      if (order == 1) { if ((rc = launch_fast())) return rc; }
    }
    if (timing) cudaEventRecord(ev[2], st);
    rc = stage_sharded_impl(Pb, Qb, accPb, accQb, GQb, HQb, nranks, rank, d, S, B, lr, reg, reg_adv, eps, adver, ws,
                            ws_bytes, stats, s, 1, err, st);
    if (rc) return rc;
    if (timing) cudaEventRecord(ev[3], st);
    if ((rc = barrier())) return rc;
    if (timing) cudaEventRecord(ev[4], st);
    rc = stage_sharded_impl(Pb, Qb, accPb, accQb, GQb, HQb, nranks, rank, d, S, B, lr, reg, reg_adv, eps, adver, ws,
                            ws_bytes, stats, s, 2, err, st);
    if (rc) return rc;
    if (timing) cudaEventRecord(ev[5], st);
    if (order == 2 && adver) { if ((rc = launch_fast())) return rc; }
    APR_CUDA_CHECK(cudaStreamWaitEvent(st, ax.join, 0));
    if (pairs_enabled(d)) APR_CUDA_CHECK(cudaStreamWaitEvent(st, ax.join2, 0));
    if ((rc = barrier())) return rc;
    if (timing && adver) {
      cudaStreamSynchronize(st);
      float t;
      cudaEventElapsedTime(&t, ev[0], ev[1]); acc_ms[0] += t;   // stage 0
      cudaEventElapsedTime(&t, ev[1], ev[2]); acc_ms[1] += t;   // barrier
      cudaEventElapsedTime(&t, ev[2], ev[3]); acc_ms[2] += t;   // stage 1
      cudaEventElapsedTime(&t, ev[3], ev[4]); acc_ms[3] += t;   // barrier
      cudaEventElapsedTime(&t, ev[4], ev[5]); acc_ms[4] += t;   // stage 2
      cudaEventElapsedTime(&t, ev[6], ev[7]); acc_ms[5] += t;   // fast kernel (second stream)
      cudaEventElapsedTime(&t, ev[0], ev[5]); acc_ms[6] += t;   // general path total
    }
  }
  if (timing) {
    if (rank == 0 && adver)
      fprintf(stderr, "[shard timing] per step (us): stage0 %.1f  bar %.1f  stage1 %.1f  bar %.1f  stage2 %.1f | fast %.1f | general total %.1f\n",
              1e3 * acc_ms[0] / count, 1e3 * acc_ms[1] / count, 1e3 * acc_ms[2] / count, 1e3 * acc_ms[3] / count,
              1e3 * acc_ms[4] / count, 1e3 * acc_ms[5] / count, 1e3 * acc_ms[6] / count);
    for (auto& e : ev) cudaEventDestroy(e);
  }
  return APR_OK;
}

/* Synchronises the stream and returns the workspace's sticky status word: bit 0 = an id outside its table reached the
 * index preparation since apr_train_workspace_init (such a triple trained row 0 instead of faulting). */
int apr_train_status(const void* ws, int32_t* flags_host, apr_stream_t stream) {
  if (!ws || !flags_host) return APR_E_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  APR_CUDA_CHECK(cudaMemcpyAsync(flags_host, ws, 4, cudaMemcpyDeviceToHost, st));
  APR_CUDA_CHECK(cudaStreamSynchronize(st));
  return APR_OK;
}

int apr_train_unique_counts(const void* ws, int64_t ws_bytes, int32_t S, int32_t B, int32_t d, int32_t* counts_host,
                            apr_stream_t stream) {
  if (!ws || !counts_host || S < 1 || B < 1 || !valid_dim(d)) return APR_E_ARG;
  TrainLayout L;
  if (!layout_for_ws(ws_bytes, B, d, S, L)) return APR_E_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int32_t* tmp = static_cast<int32_t*>(malloc(size_t(S) * 8 + 8));
  if (!tmp) return APR_E_ARG;
  const char* base = static_cast<const char*>(ws);
  cudaError_t e = cudaMemcpyAsync(tmp, base + L.off_ucnt, size_t(S) * 4, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(tmp + S, base + L.off_iall, size_t(S) * 4, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(tmp + 2 * S, base + L.off_hdr, 4, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) { free(tmp); set_cuda_error(e, "apr_train_unique_counts"); return APR_E_CUDA; }
  for (int s = 0; s < S; ++s) { counts_host[2 * s] = tmp[s]; counts_host[2 * s + 1] = tmp[S + s]; }
  const int32_t flag = tmp[2 * S];
  free(tmp);
  return flag ? APR_E_ARG : APR_OK;
}

int apr_loss_acc(const float* P, const float* Q, int32_t d, const int32_t* u, const int32_t* i, const int32_t* j, int32_t S,
                 int32_t B, double* out, apr_stream_t stream) {
  if (!P || !Q || !u || !i || !j || !out || S < 1 || B < 1 || !valid_dim(d)) return APR_E_ARG;
  if (!aligned16(P) || !aligned16(Q)) return APR_E_ALIGN;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  APR_CUDA_CHECK(cudaMemsetAsync(out, 0, size_t(S) * 2 * sizeof(double), st));
  const int q = d / 4;
  if (q <= 4) return run_loss_acc<4, 1>(P, Q, d, u, i, j, S, B, out, st);
  if (q <= 8) return run_loss_acc<8, 1>(P, Q, d, u, i, j, S, B, out, st);
  if (q <= 16) return run_loss_acc<16, 1>(P, Q, d, u, i, j, S, B, out, st);
  if (q <= 32) return run_loss_acc<32, 1>(P, Q, d, u, i, j, S, B, out, st);
  if (q <= 64) return run_loss_acc<32, 2>(P, Q, d, u, i, j, S, B, out, st);
  if (q <= 96) return run_loss_acc<32, 3>(P, Q, d, u, i, j, S, B, out, st);
  return run_loss_acc<32, 4>(P, Q, d, u, i, j, S, B, out, st);
}

}  // extern "C"
