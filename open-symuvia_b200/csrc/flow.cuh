// flow.cuh — the car-following update in the reference's leader-first order (rows A4/A5 of SURVEY.md §8a), without recursion.
//
// The reference computes a vehicle's leader before the vehicle itself (Vehicule::CalculTraficEx recursion,
// vehicule.cpp:1304-1309) because Newell's congested branch uses the leader's NEW speed
// (NewellCarFollowing.cpp:213-216).  The leader of every vehicle is fixed before the update (phase A), so the new speeds are
// the unique solution of  v0[x] = B_x(v0[leader(x)])  on the leader graph — a forest once leader loops are cut — and any
// evaluation order that ends in a state where every vehicle has been computed from its leader's FINAL speed gives the
// reference's bits.  The device gets there by chaotic relaxation instead of walking the chains:
//   * k_seg      cuts every lane (already ordered by position) into SEGMENTS: maximal runs in which each vehicle follows
//                the vehicle directly ahead.  Only a segment's front vehicle can depend on a vehicle outside it, so leader
//                loops ("Infinite loop detected", vehicule.cpp:1201-1216) are loops of the much smaller segment graph.
//   * k_cyc_*    a short walk per segment, then pointer jumping (4 steps per round) for the chains still open, finds every loop
//                and its basin on the segment graph; the basin's first
//                vehicle in m_LstVehicles order is the root the reference's recursion would start from, which fixes WHERE
//                the loop is cut: the re-entered vehicle X is computed last, its follower Y uses the estimated leader speed
//                (NewellCarFollowing.cpp:217-230).  After the cuts the leader graph is acyclic.
//   * k_relax    one thread per vehicle, all vehicles at once: phase B from the row and the leader's current speed
//                (initial guess: its start-of-step speed); a sweep recomputes only the vehicles whose leader's speed changed
//                bit-wise since they last used it.  A sweep that changes nothing is the fixed point; on an acyclic graph it
//                is unique and is reached after at most depth+1 sweeps — in practice a handful, because the follower's
//                sensitivity to the leader's new speed is dt - deltaN*tau (NewellCarFollowing.cpp:361-364), ~0 in steady
//                car following, so a perturbation drops below one ulp within two or three vehicles.
#pragma once
#include "micro.cuh"
#include "rows.cuh"

namespace symb200 {

constexpr int kRelaxBatch = 8;   // sweeps launched between two looks at the change counters

struct FlowCtl { int n_segments, n_unres, converged; unsigned gbar; int changed[kRelaxBatch + 1]; int pad3[3]; };

struct DevFlow {
  int *order, *posidx, *isfront, *seg_len, *seg_dep, *segpos, *minslot, *cut, *fronts, *ulist;
  FlowCtl* ctl;
  // loop detection scratch (indexed by order position)
  int *A, *A2, *Mn, *Mn2, *oncyc, *cid;
  unsigned long long* rootkey;
};

// `by_id`: m_LstVehicles order is the slot order on one GPU; in a partitioned run slots are local, so the vehicle id is the key
// (pre-seeded vehicles are listed in id order, reseau.cpp:1293-1340).
__device__ __forceinline__ int list_key(const DevVeh& V, int x, int by_id) { return by_id ? V.id[x] : x; }
// One thread per position of the lane order.  A position is a segment FRONT when it is the last of its lane or its vehicle does not
// follow the vehicle directly ahead; the front's thread then walks its segment towards the rear (a handful of vehicles), gives every
// member the segment's position and finds the member that comes first in m_LstVehicles (the loop-cutting rule needs it).
__global__ void k_seg(DevVeh V, DevFlow F, int n_sorted, const int* slane, const int* lane_count, const int* lane_start, int by_id) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  bool front = false;
  if (p < n_sorted) {
    const int ln = slane[p]; const int s = lane_start[ln]; const int last = s + lane_count[ln] - 1;
    const int x = F.order[p];
    int d = V.ctx_leader[x];
    front = p == last || d != F.order[p + 1];
    if (front) {
      if (d >= 0 && V.status[d] == ST_GHOST) d = -1;          // leader on another GPU: estimated speed, no dependency
      F.seg_dep[p] = d;
      int mn = list_key(V, x, by_id), mnx = x;
      F.segpos[x] = p;
      for (int q = p - 1; q >= s; q--) {                      // members: positions below the front whose vehicle follows the one directly ahead
        const int y = F.order[q];
        if (V.ctx_leader[y] != F.order[q + 1]) break;
        F.segpos[y] = p;
        const int k = list_key(V, y, by_id); if (k < mn) { mn = k; mnx = y; }
      }
      F.minslot[p] = mnx;
    }
  }
  // compact list of the segment fronts (any order): one atomic per warp
  const unsigned m = __ballot_sync(0xffffffffu, front);
  if (m) { const int lane = threadIdx.x & 31; int base = 0;
    if (lane == 0) base = atomicAdd(&F.ctl->n_segments, __popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (front) F.fronts[base + __popc(m & ((1u << lane) - 1))] = p; }
}

// one relaxation sweep.  first: every vehicle, and the dependency (order position -> leader slot, -1 = none / cut / ghost)
// is cached; later sweeps only look at the leader's speed again.  `prev` = changes counted by the sweep before.
// Leaders are mostly the neighbour in the lane order, i.e. in the same warp: every warp repeats its part until
// nothing in it changes (kLocalIters at most), so a global sweep carries a change down whole segments, not one vehicle.
#ifndef SYM_LOCAL_ITERS
#define SYM_LOCAL_ITERS 48
#endif
constexpr int kLocalIters = SYM_LOCAL_ITERS;
#ifndef SYM_RELAX_BLOCKS
#define SYM_RELAX_BLOCKS 8   // 128 threads x 8 blocks per SM, 64 registers (C4, ms per step: 256x3 0.38, 256x5 0.32, 128x10 0.30, 128x8 0.297, 512x2 0.35)
#endif
#ifndef SYM_RELAX_THREADS
#define SYM_RELAX_THREADS 128
#endif
__global__ void __launch_bounds__(SYM_RELAX_THREADS, SYM_RELAX_BLOCKS) k_relax(DevNet N, DevVeh V, DevLaneDyn LD, const double* __restrict__ rows, double* b_goal, int* b_flags,
                                               double* vl_used, int* rl_lead, int n_sorted, int first, const int* prev, int* changed, double inst, const DevMrg* M, const double* mrgx) {
  if (!first && *prev == 0) return;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const Row r = row_at(rows, p < n_sorted ? p : 0);
  int dep = -1; double vL = 0, vl_last = 0; bool need = false, cutl = false, touched = false;
  if (p < n_sorted) {
    if (first) {
      const int L = lo32(r[R_I0]); cutl = hi32(r[R_I1]) & 1;
      dep = (L >= 0 && !cutl) ? L : -1; rl_lead[p] = dep;
      if (L >= 0) vL = cutl ? r[R_VLEST] : __ldcg(&V.vit_n[L]);
      need = true;
    } else {
      dep = rl_lead[p];
      if (dep >= 0) { vL = __ldcg(&V.vit_n[dep]); vl_last = vl_used[p]; need = __double_as_longlong(vL) != __double_as_longlong(vl_last); }
      cutl = false;                                  // a cut leader is no dependency (dep = -1): later sweeps never recompute such a vehicle
    }
  }
  int nch = 0;
  for (int it = 0;; it++) {
    bool ch = false;
    if (need) {
      const int hdr = hi32(r[R_I0]); const int x = lo32(r[R_I1]);
      vl_last = vL; touched = true;
      double v0;
      if (hdr & H_SLOW) {                           // context longer than the row holds: general path (idempotent form)
        follow_one(cP_ref(), N, V, LD, x, !cutl, inst, r[R_CPOS], true, dep >= 0 ? &vL : nullptr, &v0, M, p, hdr);
        b_flags[p] = B_SLOWDONE;
      } else {
        BOut o = phase_b(cP_ref(), r, vL, inst, mrgx + (size_t)p * 6, cutl);
        b_goal[p] = o.goal; b_flags[p] = o.flags; V.dn[x] = o.dn_end; V.cpos[x] = o.cpos; v0 = o.v0;
      }
      if (first && it == 0) V.computed[x] = 1;
      if (__double_as_longlong(v0) != __double_as_longlong(V.vit_n[x])) { V.vit_n[x] = v0; ch = true; }
    }
    nch += ch;
#ifdef SYM_RELAX_BLOCKLOOP
    if (!__syncthreads_or(ch) || it == kLocalIters) break;             // whole block in step (C4: relax 0.316 ms)
#else
    if (!__any_sync(0xffffffffu, ch) || it == kLocalIters) break;     // every warp on its own: no barrier, other warps' changes arrive through L2 (0.293 ms)
#endif
    need = false;
    if (dep >= 0) { vL = __ldcg(&V.vit_n[dep]); need = __double_as_longlong(vL) != __double_as_longlong(vl_last); }
  }
  if (touched) vl_used[p] = vl_last;
  const unsigned m = __ballot_sync(0xffffffffu, nch > 0);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(changed, __popc(m));
}
// phase C for every vehicle of a lane (parallel)
// verdict of a batch of sweeps: did one of them change nothing?
__global__ void k_relax_done(FlowCtl* ctl, int nb) {
  if (threadIdx.x || blockIdx.x) return;
  int ok = 0; for (int j = 1; j <= nb; j++) if (ctl->changed[j] == 0) ok = 1;
  ctl->converged = ok;
}
__global__ void k_place(DevNet N, DevVeh V, const double* rows, const double* b_goal, const int* b_flags, int n_sorted, double inst, const int* guard, const DevMrg* M) {
  if (guard && *guard == 0) return;              // launched ahead of the convergence check (engine.cu: step_back)
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_sorted) return;
  phase_c(cP_ref(), N, V, rows, b_goal, b_flags, p, inst, M);
}

// ---- leader loops: pointer jumping on the segment graph ---------------------------------------------------------------
// Most chains end within a few segments (a vehicle without leader, a cut, a ghost): every front first walks kCycWalk steps on
// its own; the fronts that have not reached a terminal by then (long chains, loops and their basins) go to a short list and
// only they take part in the pointer jumping.  A[p] of a resolved front points straight at its terminal (in both buffers).
constexpr int kCycWalk = 8;
__global__ void k_cyc_walk(DevFlow F) {
  const int ns = F.ctl->n_segments;
  for (int t0 = blockIdx.x * blockDim.x; t0 < ns; t0 += gridDim.x * blockDim.x) {      // whole warps stay together for the ballot
    const int t = t0 + threadIdx.x;
    bool unres = false; int p = -1;
    if (t < ns) {
      p = F.fronts[t];
      int q = p, mn = p; bool term = false;
      for (int h = 0; h < kCycWalk; h++) { int d = F.seg_dep[q]; if (d < 0) { term = true; break; } q = F.segpos[d]; if (h + 1 < kCycWalk && q < mn) mn = q; }
      F.A[p] = q; F.A2[p] = q; F.Mn[p] = mn; F.Mn2[p] = mn;
      unres = !term;
      if (unres) { F.oncyc[p] = 0; F.rootkey[p] = ~0ULL; F.cid[p] = -1; }
    }
    const unsigned m = __ballot_sync(0xffffffffu, unres);
    if (m) { const int lane = threadIdx.x & 31; int base = 0;
      if (lane == 0) base = atomicAdd(&F.ctl->n_unres, __popc(m));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (unres) F.ulist[base + __popc(m & ((1u << lane) - 1))] = p; }
  }
}
__global__ void k_cyc_jump(DevFlow F, int flip) {  // A <- A^4, Mn <- min over the segments stepped over
  const int ns = F.ctl->n_unres;
  const int* A = flip ? F.A2 : F.A; const int* M = flip ? F.Mn2 : F.Mn;
  int* Ao = flip ? F.A : F.A2; int* Mo = flip ? F.Mn : F.Mn2;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < ns; t += gridDim.x * blockDim.x) {
    int p = F.ulist[t];
    int a = A[p], a1 = A[a], a2 = A[a1];
    Ao[p] = A[a2];
    Mo[p] = min(min(M[p], M[a]), min(M[a1], M[a2]));
  }
}
__global__ void k_cyc_mark(DevVeh V, DevFlow F, int flip, int by_id) {
  const int ns = F.ctl->n_unres;
  const int* A = flip ? F.A2 : F.A; const int* M = flip ? F.Mn2 : F.Mn;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < ns; t += gridDim.x * blockDim.x) {
    int p = F.ulist[t];
    int c0 = A[p];                                 // where this segment's chain ends up: a terminal, or a segment on a loop
    if (F.seg_dep[c0] < 0) continue;
    F.oncyc[c0] = 1;
    int id = M[c0];                                // smallest segment of that loop = its representative
    F.cid[p] = id;
    unsigned long long key = ((unsigned long long)(unsigned)list_key(V, F.minslot[p], by_id) << 32) | (unsigned)p;
    atomicMin(&F.rootkey[id], key);
  }
}
// one thread per loop: cut it where the reference's recursion would
__global__ void k_cyc_break(DevVeh V, DevFlow F, double* rows, long long* counters) {
  const int ns = F.ctl->n_unres;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < ns; t += gridDim.x * blockDim.x) {
  int p = F.ulist[t];
  if (!F.oncyc[p] || F.cid[p] != p) continue;      // representative = smallest segment of the loop
  unsigned long long key = F.rootkey[p];
  int rs = (int)(key & 0xffffffffu); int r = F.minslot[rs];
  // entry segment of the root's chain into the loop, and the vehicle it enters at
  int e = rs, T = -1;
  while (!F.oncyc[e]) { T = F.seg_dep[e]; e = F.segpos[T]; }
  // predecessor of e on the loop
  int q = e;
  for (;;) { int nx = F.segpos[F.seg_dep[q]]; if (nx == e) break; q = nx; }
  int Tloop = F.seg_dep[q];                        // vehicle at which the loop enters e (tail of e's lane)
  int X, Y;
  int cand = (e != rs) ? T : r;                   // first vehicle of the root's chain inside e
  X = (F.posidx[cand] >= F.posidx[Tloop]) ? cand : Tloop;   // behind the loop's entry point: the chain joins the loop there
  if (X == Tloop) Y = F.order[q];                  // front vehicle of the predecessor segment
  else Y = F.order[F.posidx[X] - 1];               // the vehicle right behind X inside e
  F.cut[Y] = 1;
  if (cP_ref().cls[V.cls[Y]].law == 0) atomicAdd((unsigned long long*)&counters[SYMGPU_CNT_LOOPS], 1ULL);
  int py = F.posidx[Y];
  ((int*)&row_at(rows, py)[R_I1])[1] |= 1;          // row header: leader stays uncomputed (little endian: high int)
  }
}

// The pointer-jumping rounds, the marking and the cutting in ONE launch (cooperative: all blocks resident, barriers between the
// phases on a counter in FlowCtl): the open chains are few, so the separate launches (K + 2 of them) only cost their latency.
// Data another block wrote in an earlier phase is read through L2 (__ldcg): the L1 of this SM may hold the line from before.
__device__ __forceinline__ void grid_barrier(unsigned* count, unsigned nblocks, unsigned& phase) {
  __syncthreads();
  if (threadIdx.x == 0) {
    phase++;
    __threadfence();
    atomicAdd(count, 1u);
    const unsigned target = phase * nblocks;
    while (*(volatile unsigned*)count < target) { }
    __threadfence();
  }
  __syncthreads();
}
__global__ void __launch_bounds__(256) k_cyc_rest(DevVeh V, DevFlow F, double* rows, long long* counters, int klev, int by_id) {
  unsigned phase = 0; const unsigned nb = gridDim.x;
  const int ns = __ldcg(&F.ctl->n_unres);
  if (ns == 0) return;                                   // every block sees the same count: no chain left open
  // an open chain only runs through open fronts (a front that ends within kCycWalk steps is resolved, and so is everything that reaches it
  // within the walk), so it is at most ns + kCycWalk segments long: 4^k * walk >= that is enough, whatever the host sized klev for
  { int kl = 1; while ((1LL << (2 * kl)) * kCycWalk < (long long)ns + kCycWalk) kl++; kl++; if (kl < klev) klev = kl; }
  int flip = 0;
  for (int k = 0; k < klev; k++) {                     // A <- A^4, Mn <- min over the segments stepped over
    const int* A = flip ? F.A2 : F.A; const int* M = flip ? F.Mn2 : F.Mn;
    int* Ao = flip ? F.A : F.A2; int* Mo = flip ? F.Mn : F.Mn2;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < ns; t += gridDim.x * blockDim.x) {
      int p = F.ulist[t];
      int a = __ldcg(&A[p]), a1 = __ldcg(&A[a]), a2 = __ldcg(&A[a1]);
      Ao[p] = __ldcg(&A[a2]);
      Mo[p] = min(min(__ldcg(&M[p]), __ldcg(&M[a])), min(__ldcg(&M[a1]), __ldcg(&M[a2])));
    }
    flip ^= 1;
    grid_barrier(&F.ctl->gbar, nb, phase);
  }
  { const int* A = flip ? F.A2 : F.A; const int* M = flip ? F.Mn2 : F.Mn;          // mark (k_cyc_mark)
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < ns; t += gridDim.x * blockDim.x) {
      int p = F.ulist[t];
      int c0 = __ldcg(&A[p]);
      if (F.seg_dep[c0] < 0) continue;
      F.oncyc[c0] = 1;
      int id = __ldcg(&M[c0]);
      F.cid[p] = id;
      unsigned long long key = ((unsigned long long)(unsigned)list_key(V, F.minslot[p], by_id) << 32) | (unsigned)p;
      atomicMin(&F.rootkey[id], key);
    } }
  grid_barrier(&F.ctl->gbar, nb, phase);
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < ns; t += gridDim.x * blockDim.x) {   // cut (k_cyc_break)
    int p = F.ulist[t];
    if (!__ldcg(&F.oncyc[p]) || __ldcg(&F.cid[p]) != p) continue;
    unsigned long long key = __ldcg(&F.rootkey[p]);
    int rs = (int)(key & 0xffffffffu); int r = F.minslot[rs];
    int e = rs, T = -1;
    while (!__ldcg(&F.oncyc[e])) { T = F.seg_dep[e]; e = F.segpos[T]; }
    int q = e;
    for (;;) { int nx = F.segpos[F.seg_dep[q]]; if (nx == e) break; q = nx; }
    int Tloop = F.seg_dep[q];
    int cand = (e != rs) ? T : r;
    int X = (F.posidx[cand] >= F.posidx[Tloop]) ? cand : Tloop;
    int Y = (X == Tloop) ? F.order[q] : F.order[F.posidx[X] - 1];
    F.cut[Y] = 1;
    if (cP_ref().cls[V.cls[Y]].law == 0) atomicAdd((unsigned long long*)&counters[SYMGPU_CNT_LOOPS], 1ULL);
    int py = F.posidx[Y];
    ((int*)&row_at(rows, py)[R_I1])[1] |= 1;
  }
}

}  // namespace symb200
