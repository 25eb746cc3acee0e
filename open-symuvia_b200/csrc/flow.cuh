// flow.cuh — dependency-ordered execution of the car-following update (rows A4/A5 of SURVEY.md §8a).
//
// The reference computes a vehicle's leader before the vehicle itself (Vehicule::CalculTraficEx recursion,
// vehicule.cpp:1304-1309) because Newell's congested branch uses the leader's NEW speed
// (NewellCarFollowing.cpp:213-216).  On the device the same order is obtained without recursion:
//   * k_seg      cuts every lane (already ordered by position) into SEGMENTS: maximal runs in which each
//                vehicle follows the vehicle directly ahead.  Only a segment's front vehicle can depend on
//                a vehicle outside the segment (the tail of the next lane of its route).
//   * k_flow     persistent dataflow kernel: one warp = one worker; ready segments are taken from a ticket
//                queue, walked front -> back by lane 0, and on completion the segments waiting on them are
//                released (first one handed to the same worker: no queue latency on the critical chain).
//                No worker ever waits on unscheduled work: the only spin is on the queue, and it ends when
//                every scheduled segment has completed.
//   * k_cyc_*    when segments remain (a loop of leaders, "Infinite loop detected" vehicule.cpp:1201-1216):
//                pointer-doubling on the segment graph finds every loop and its basin; the basin's first
//                vehicle in m_LstVehicles order is the root the reference's recursion would start from, which
//                fixes WHERE the loop is cut (the re-entered vehicle X stays uncomputed, its follower Y uses
//                the estimated leader speed, NewellCarFollowing.cpp:217-230); then the flow restarts.
#pragma once
#include "micro.cuh"
#include "rows.cuh"

namespace symb200 {

constexpr int kDone = -2;     // first_waiter value of a completed segment

struct FlowCtl { int q_head, n_pushed, n_started, n_completed, n_segments, n_vehicles_done, pad0, pad1; };

struct DevFlow {
  int *order, *posidx, *isfront, *seg_len, *seg_dep, *segpos, *minslot, *first_waiter, *next_waiter, *cut, *queue;
  FlowCtl* ctl;
  // loop detection scratch (indexed by order position)
  int *A, *A2, *Mn, *Mn2, *oncyc, *cid;
  unsigned long long* rootkey;
};

// per lane: segments, front -> rear
// `by_id`: m_LstVehicles order is the slot order on one GPU; in a partitioned run slots are local, so the vehicle id is the key
// (pre-seeded vehicles are listed in id order, reseau.cpp:1293-1340).
__device__ __forceinline__ int list_key(const DevVeh& V, int x, int by_id) { return by_id ? V.id[x] : x; }
__global__ void k_seg(DevVeh V, DevFlow F, int n_lanes, const int* lane_count, const int* lane_start, int by_id) {
  int ln = blockIdx.x * blockDim.x + threadIdx.x;
  if (ln >= n_lanes) return;
  int n = lane_count[ln];
  if (!n) return;
  int s = lane_start[ln];
  int cur = -1, nseg = 0, len = 0, mn = 0x7fffffff, mnx = -1;
  for (int a = n - 1; a >= 0; a--) {
    int p = s + a; int x = F.order[p];
    int d = V.ctx_leader[x];
    bool front = (a == n - 1) || d != F.order[p + 1];
    if (d >= 0 && V.status[d] == ST_GHOST) d = -1;          // leader on another GPU: estimated speed, no dependency
    if (front) {
      if (cur >= 0) { F.seg_len[cur] = len; F.minslot[cur] = mnx; }
      cur = p; len = 0; mn = 0x7fffffff; mnx = -1; nseg++;
      F.isfront[p] = 1; F.seg_dep[p] = d; F.first_waiter[p] = -1;
    } else F.isfront[p] = 0;
    F.segpos[x] = cur; F.cut[x] = 0;
    len++; { int k = list_key(V, x, by_id); if (k < mn) { mn = k; mnx = x; } }
  }
  F.seg_len[cur] = len; F.minslot[cur] = mnx;
  atomicAdd(&F.ctl->n_segments, nseg);
}

__device__ __forceinline__ void flow_push(const DevFlow& F, int seg) {
  atomicAdd(&F.ctl->n_started, 1);
  int slot = atomicAdd(&F.ctl->n_pushed, 1);
  __threadfence();
  *(volatile int*)&F.queue[slot] = seg;
}

// start of a round: forget the waiter lists of unfinished segments
__global__ void k_flow_reset(DevFlow F, int n_sorted) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_sorted) return;
  if (F.isfront[p] && F.first_waiter[p] != kDone) F.first_waiter[p] = -1;
}
// schedule the segments that can start, register the others on the segment they wait for
__global__ void k_flow_init(DevVeh V, DevFlow F, int n_sorted) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_sorted) return;
  if (!F.isfront[p] || F.first_waiter[p] == kDone) return;
  int x = F.order[p];
  int d = F.seg_dep[p];
  if (d < 0 || F.cut[x] || V.computed[d]) { flow_push(F, p); return; }
  int t = F.segpos[d];
  for (;;) {
    int h = *(volatile int*)&F.first_waiter[t];
    if (h == kDone) { flow_push(F, p); return; }
    F.next_waiter[p] = h;
    __threadfence();
    if (atomicCAS(&F.first_waiter[t], h, p) == h) return;
  }
}

// persistent workers.  Rows of a segment are contiguous (descending order position): the warp stages them in shared
// memory with coalesced 256-byte loads and lane 0 runs phase B down the chain out of shared memory.  While lane 0
// computes, lanes 1..31 already stage the rows of the first segment waiting on this one (the waiter lists are frozen
// once k_flow_init has run), which the same worker continues with: along a dependency chain no queue, fence or
// global-memory round trip sits between two segments.
constexpr int kBufRows = 16;
__global__ void __launch_bounds__(128, 3) k_flow(DevNet N, DevVeh V, DevLaneDyn LD, DevFlow F, const double* __restrict__ rows,
                                              double* b_goal, int* b_flags, double inst) {
  __shared__ double srow[4][2][kBufRows][kRowD];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int seg = -1, len = 0, buf = 0;
  bool loaded = false;
  for (;;) {
    if (seg < 0) {
      if (lane == 0) {                             // take a ticket
        int t = atomicAdd(&F.ctl->q_head, 1);
        for (;;) {
          int v = *(volatile int*)&F.queue[t];
          if (v >= 0) { seg = v; break; }
          int c = *(volatile int*)&F.ctl->n_completed;
          int st = *(volatile int*)&F.ctl->n_started;
          int q = *(volatile int*)&F.ctl->n_pushed;
          if (c == st && t >= q) { seg = -1; break; }   // everything scheduled has completed: no work will ever come
          __nanosleep(32);
        }
      }
      seg = __shfl_sync(0xffffffffu, seg, 0);
      if (seg < 0) return;
      loaded = false;
    }
    if (!loaded) {
      len = F.seg_len[seg];
      const int cnt = min(len, kBufRows);
      for (int j = 0; j < cnt; j++) srow[w][buf][j][lane] = __ldcg(&rows[(size_t)(seg - j) * kRowD + lane]);
      __syncwarp();
    }
    int h = -1, hlen = 0;
    if (lane != 0) {                               // stage the successor
      h = F.first_waiter[seg];
      if (h >= 0) {
        hlen = F.seg_len[h];
        const int cnt = min(hlen, kBufRows);
        for (int j = 0; j < cnt; j++) {
          srow[w][buf ^ 1][j][lane] = __ldcg(&rows[(size_t)(h - j) * kRowD + lane]);
          if (lane == 1) srow[w][buf ^ 1][j][0] = __ldcg(&rows[(size_t)(h - j) * kRowD]);
        }
      }
    } else {                                       // phase B down the segment
      int prev_x = -1; double prev_v0 = 0;
      for (int j = 0; j < len; j++) {
        const int p = seg - j;
        const double* r = j < kBufRows ? srow[w][buf][j] : rows + (size_t)p * kRowD;
        const int hdr = hi32(r[R_I0]); const int x = lo32(r[R_I1]);
        const bool uncomputed_leader = j == 0 && (hi32(r[R_I1]) & 1);
        if (hdr & H_SLOW) {
          follow_one(cP_ref(), N, V, LD, x, !uncomputed_leader, inst);
          b_flags[p] = B_SLOWDONE; prev_x = x; prev_v0 = V.vit_n[x];
        } else {
          const int L = lo32(r[R_I0]);
          double vL = 0;
          if (L >= 0) vL = uncomputed_leader ? r[R_VLEST] : (L == prev_x ? prev_v0 : __ldcg(&V.vit_n[L]));
          BOut o = phase_b(cP_ref(), r, vL, inst);
          V.vit_n[x] = o.v0; b_goal[p] = o.goal; b_flags[p] = o.flags; V.dn[x] = o.dn_end; V.cpos[x] = o.cpos;
          prev_x = x; prev_v0 = o.v0;
        }
        V.computed[x] = 1;
      }
    }
    __syncwarp();
    h = __shfl_sync(0xffffffffu, h, 1); hlen = __shfl_sync(0xffffffffu, hlen, 1);
    if (lane == 0) {
      if (h >= 0) {
        atomicAdd(&F.ctl->n_started, 1);           // the first waiter continues on this worker
        int o = F.next_waiter[h];
        while (o >= 0) { int nx = F.next_waiter[o]; flow_push(F, o); o = nx; }   // the others go through the queue (release inside)
      }
      F.first_waiter[seg] = kDone;
      atomicAdd(&F.ctl->n_vehicles_done, len);
      atomicAdd(&F.ctl->n_completed, 1);
    }
    seg = h; len = hlen; buf ^= 1; loaded = h >= 0;
  }
}
// phase C for every vehicle of a lane (parallel)
__global__ void k_place(DevNet N, DevVeh V, const double* rows, const double* b_goal, const int* b_flags, int n_sorted, double inst) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_sorted) return;
  phase_c(cP_ref(), N, V, rows, b_goal, b_flags, p, inst);
}

// ---- loop detection on the remaining segments ----------------------------------------------------------------
__global__ void k_cyc_init(DevFlow F, int n_sorted) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_sorted) return;
  F.oncyc[p] = 0; F.rootkey[p] = ~0ULL; F.cid[p] = -1;
  if (!F.isfront[p] || F.first_waiter[p] == kDone) { F.A[p] = p; F.Mn[p] = p; return; }
  F.A[p] = F.segpos[F.seg_dep[p]];                 // every remaining segment waits on a remaining segment
  F.Mn[p] = p;
}
__global__ void k_cyc_double(DevFlow F, int n_sorted, int flip) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_sorted) return;
  const int* A = flip ? F.A2 : F.A; const int* M = flip ? F.Mn2 : F.Mn;
  int* Ao = flip ? F.A : F.A2; int* Mo = flip ? F.Mn : F.Mn2;
  int a = A[p];
  Ao[p] = A[a];
  int m = M[p], m2 = M[a];
  Mo[p] = m2 < m ? m2 : m;
}
__global__ void k_cyc_mark(DevVeh V, DevFlow F, int n_sorted, int flip, int by_id) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_sorted) return;
  if (!F.isfront[p] || F.first_waiter[p] == kDone) return;
  const int* A = flip ? F.A2 : F.A; const int* M = flip ? F.Mn2 : F.Mn;
  int c0 = A[p];                                   // a segment on the loop this one leads to
  F.oncyc[c0] = 1;
  int id = M[c0];                                  // smallest segment id of that loop
  F.cid[p] = id;
  unsigned long long key = ((unsigned long long)(unsigned)list_key(V, F.minslot[p], by_id) << 32) | (unsigned)p;
  atomicMin(&F.rootkey[id], key);
}
// one thread per loop: cut it where the reference's recursion would
__global__ void k_cyc_break(DevVeh V, DevFlow F, double* rows, int n_sorted, long long* counters) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_sorted) return;
  if (!F.oncyc[p] || F.cid[p] != p) return;        // representative = smallest segment of the loop
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
  ((int*)&rows[(size_t)py * kRowD + R_I1])[1] |= 1;   // row header: leader stays uncomputed (little endian: high int)
  if (!F.isfront[py]) {                            // split e below X: [rear .. Y] becomes a segment of its own
    int rear = e - F.seg_len[e] + 1;
    F.isfront[py] = 1; F.seg_len[py] = py - rear + 1; F.seg_len[e] = e - py; F.seg_dep[py] = X; F.first_waiter[py] = -1;
    for (int a = rear; a <= py; a++) F.segpos[F.order[a]] = py;
    atomicAdd(&F.ctl->n_segments, 1);
  }
}

}  // namespace symb200
