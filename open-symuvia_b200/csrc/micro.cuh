// micro.cuh — per-vehicle device code of the microscopic update (sm_100a).
//
// Struct-of-arrays restatement of the per-vehicle part of SymuVia's time step.  Each device
// function cites the reference code it replaces (paths relative to symuvia/src of
// licit-lab/Open-SymuVia).  Arithmetic is IEEE fp64 with FMA contraction disabled (-fmad=false)
// and operations written in the reference's order, so results can be compared bit for bit
// with a non-FMA x86 build of the reference.
#pragma once
#include <cfloat>
#include <cmath>
#include <cstdint>

namespace symb200 {

#define SYM_UNDEF (-DBL_MIN)     /* symUtil.h:9 UNDEF_POSITION */

constexpr int kMaxCtx = 8;       // longest context lane list (CarFollowingContext::BuildLstLanes) handled on device
constexpr int kMemo = 4;         // per-vehicle next-lane draw memo entries (Vehicule::m_NextVoieTirage)
constexpr int kMaxCls = 8;

// vehicle slot status
enum : int { ST_DEAD = 0, ST_ALIVE = 1, ST_NEW = 2 /* entered m_LstVehicles this step, never computed */,
             ST_POOL = 3 /* head of an origin's waiting queue, not in m_LstVehicles */,
             ST_GHOST = 4 /* read-only copy of the tail vehicle of a lane owned by another GPU (partitioned runs) */ };
// persistent per-vehicle flags
enum : int { F_HASCTX = 1, F_PREVFREE = 2 };
// output flags (same packing as the oracle dump): bit0 regime fluide, bit1 arret au feu, bit2 context free-flow
enum : int { O_FLUIDE = 1, O_FEU = 2, O_CTXFREE = 4, O_ENDLEG = 16, O_CREATED = 32, O_LINK1NULL = 64, O_VDESRESET = 128 };

struct DevCls { double w, kx, vx, esp, decel, tau_lc, pi_rab; int law, nplage; double ax[4], vs[4]; };
struct DevParams { double dt, relax, duree, min_kv, bevel; int accbornee, ncls; DevCls cls[kMaxCls]; double maxinf[kMaxCls]; };

struct DevNet {
  const int *lane_link, *lane_num, *lane_prev, *lane_flags; const double* lane_len;
  const int *succ_off, *succ_key, *succ_next;
  const int *sl_off, *sl_i; const double* sl_pos;              // sl_i stride 6
  const int *link_first, *link_nl, *link_up, *link_down, *link_type, *link_brick; const double *link_len, *link_vreg;
  const int *route_off, *route_links;
  const int* sl_col; const double* sl_prop;                    // per stop line, refreshed every step (k_signals)
  int n_lanes, n_links, n_sl;
};

// Per-vehicle memo of the table walks of phase A that only depend on (vehicle, lane): the next lanes of the itinerary from the
// vehicle's lane (Vehicule::CalculNextVoie applied 4 times, with the lane whose draw each application consults), which of them
// are junction-internal, and for the first three context lanes THE stop line of the vehicle's movement (CalculTraficFeux,
// vehicule.cpp:1829-1883: upstream lane and downstream lane both match).  Valid while the vehicle stays on `lane`; rebuilt by
// k_ctx_fill for the vehicles that changed lane (a few per cent per step), so k_context walks no table for the others.
struct CtxCache {
  int lane, id;          // key
  int chain[4];          // chain[h] = CalculNextVoie^(h+1)(lane), -1 from the first failure on
  int tl[4];             // lane whose memoised draw hop h consults (-1 none)
  int slm[3];            // context lane k: first stop line matching the movement, -1 none, -2 downstream lane beyond the memo
  int bits;              // bit k (k=1,2): lane k has a stop line of its upstream lane at pos >= 0; bit 8+h: chain[h] is internal;
                         // bit 16+k: more than one matching stop line on lane k
  double slu0;           // lane 0: largest position of a stop line of its upstream lane (-DBL_MAX none)
};
static_assert(sizeof(CtxCache) == 64, "CtxCache is one 64-byte record");

struct DevVeh {
  int *cc_lane, *cc_id; CtxCache* cc;     // cc_lane / cc_id: the cache key again, struct-of-arrays, for the staleness test of k_scatter
  int *status, *id, *cls, *route, *route_pos, *lane, *lane_n, *prev_leader, *flags, *oflags, *memo, *vdes;
  double *pos, *vit, *acc, *pos_n, *vit_n, *acc_n, *dn, *cpos, *tcre, *dstp, *hent, *tsort;
  // per-step scratch
  int *ahead, *computed, *ctx_nl, *ctx_lane, *ctx_leader, *ctx_flags;
  double *ctx_gap, *ctx_dn0, *ctx_dfree;
};

struct DevLaneDyn { int *count, *start, *fill, *tail, *head_done; double* next_sortie; int* exit_winner; };
struct DevMrg;                                                   // merge points (merge.cuh)
struct MrgA { bool own = false, lead = false; double mc = 0, posl = 0, lenl = 0, borne = 0, repl = 0, vunc = 0; };   // phase A results for the merge hooks
__device__ inline int mrg_phase_a(const DevParams& P, const DevNet& n, const DevVeh& V, const DevMrg* M, int i, int lane1, int L, double p1, double dtv,
                                  double inst, int* memo, MrgA* out);
__device__ inline void mrg_store(const DevMrg* M, int p, const MrgA& a);
__device__ inline void mrg_load(const DevMrg* M, int p, int hdr, MrgA* a);
__device__ inline void mrg_register(const DevNet& n, const DevMrg* M, int i, int p, const int* lanes, int kf);

// std::pow with the three exponents the laws use (GippsCarFollowing.cpp:75 gamma = 0.5, IDMCarFollowing.cpp:148 alpha = 4 and
// :175 beta = 1.5), correctly rounded: the host libm the reference links (glibc) returns the correctly rounded value for all
// but ~1e-5 of the arguments, CUDA's pow() is allowed 2 ulp — and the car-following feedback amplifies one ulp beyond the
// 1e-9 gate within a few hundred steps.  Evaluated in double-double (error-free products through fma), rounded once.
__device__ __forceinline__ double pow_half(double x) { return sqrt(x); }
__device__ __forceinline__ double pow_four(double x) {
  double p1 = x * x, e1 = fma(x, x, -p1);            // x^2 = p1 + e1 exactly
  double p2 = p1 * p1, e2 = fma(p1, p1, -p2);        // p1^2 = p2 + e2 exactly
  e2 = e2 + 2.0 * p1 * e1;                           // (p1 + e1)^2 = p2 + e2 + 2 p1 e1 (+ e1^2 ~ 2^-106)
  return p2 + e2;
}
__device__ __forceinline__ double pow_three_halves(double z) {
  if (!(z > 0.0) || z > DBL_MAX) return pow(z, 1.5);
  double s = sqrt(z); double r = fma(-s, s, z);      // z = s^2 + r exactly
  double c = r / (2.0 * s);                          // sqrt(z) = s + c (1 + O(2^-52))
  double p = z * s, e = fma(z, s, -p);               // z s = p + e exactly
  e = e + z * c;
  return p + e;
}
__device__ __forceinline__ double snap_pos(double p) {           // vehicule.cpp:995-997
  return (p != SYM_UNDEF && fabs(p) < 0.00001) ? 0.0 : p;
}
__host__ __device__ __forceinline__ double acc_max(const DevCls& c, double v) {   // vehicule.cpp:157-175
  for (int i = 0; i < c.nplage; i++) if (v <= c.vs[i] + 0.000001) return c.ax[i];
  return c.nplage ? c.ax[c.nplage - 1] : 0.0;
}
__device__ __forceinline__ double len_ex(const DevParams& P, const DevNet& n, int lane) {   // voie.cpp:1451-1457
  return (n.lane_flags[lane] & 1) ? n.lane_len[lane] - P.bevel : n.lane_len[lane];
}
__host__ __device__ __forceinline__ double max_influence(const DevParams& P, const DevCls& c) {
  if (c.law == 0) return 1.0 / P.min_kv;                                  // NewellCarFollowing.cpp:48-51
  if (c.law == 1) return 3 * c.vx * P.dt;                                 // GippsCarFollowing.cpp:45-48
  double dec = c.decel; if (dec == 0) dec = DBL_MAX;                      // IDMCarFollowing.cpp:49-59
  double s = c.esp + fmax(1 * c.vx + c.vx * c.vx / (2.0 * sqrt(acc_max(c, c.vx) * dec)), 0.0);
  return 3 * s;
}

// Vehicule::GetItineraireIndex(isSameTuyau) on a repeat-free itinerary, searched forward from the cursor.
__device__ inline int route_index(const DevNet& n, int route, int rpos, int link) {
  int o = n.route_off[route], e = n.route_off[route + 1];
  for (int k = o + (rpos > 0 ? rpos : 0); k < e; k++) if (n.route_links[k] == link) return k - o;
  for (int k = o; k < e && k - o < rpos; k++) if (n.route_links[k] == link) return k - o;
  return -1;
}

// Vehicule::CalculNextVoie, 'iti' branch (vehicule.cpp:3070-3469) over the flattened movement table.
// *tirage = upstream lane whose memoised draw (Vehicule::GetTirage, :1779-1795) the call consults, or -1.
// Every (lane, next link) movement has exactly one target lane with rate 1 (checked at load), so the
// value of the draw cannot change the result; only the draw COUNT is part of the contract.
__device__ inline int next_voie(const DevNet& n, int route, int rpos, int lane, int* tirage) {
  *tirage = -1;
  if (lane < 0) return -1;
  if (n.lane_flags[lane] & 1) return -1;                                  // :3105-3108
  int link = n.lane_link[lane];
  int ta = n.link_type[link];
  if (ta == 'S' || ta == 'P' || ta == 'E') return -1;
  int ref = link, lam = lane;
  if (n.link_brick[link] >= 0) { lam = n.lane_prev[lane]; ref = n.lane_link[lam]; }   // :3172-3199
  int i = route_index(n, route, rpos, ref);                               // :3350
  int o = n.route_off[route], e = n.route_off[route + 1];
  if (i < 0 || o + i + 1 >= e) return -1;
  int ts = n.route_links[o + i + 1];
  for (int m = n.succ_off[lane]; m < n.succ_off[lane + 1]; m++)
    if (n.succ_key[m] == ts) { *tirage = lam; return n.succ_next[m]; }    // :3410-3461
  return -1;                                                              // movement not allowed (:3446-3449)
}

// memo of lanes already drawn for; returns 1 when a new draw is consumed
__device__ inline int memo_touch(int* memo, int lane) {
  for (int k = 0; k < kMemo; k++) if (memo[k] == lane) return 0;
  for (int k = kMemo - 1; k > 0; k--) memo[k] = memo[k - 1];
  memo[0] = lane;
  return 1;
}

// -------------------------------------------------------------------------------------------------
// Phase A (no dependency on other vehicles' new state): context, leader, deltaN at start of step,
// free-flow distance.  CarFollowingContext::Build (CarFollowingContext.cpp:38-170),
// NewellContext::ComputeDeltaN (NewellContext.cpp:121-169), NewellCarFollowing::ComputeFreeFlowDistance
// (NewellCarFollowing.cpp:70-111), and the CalculNextVoie side effects of Vehicule::CalculTraficFeux
// (vehicule.cpp:1854-1874).  `own_leader` = nearest vehicle ahead on the own lane (-1 none).
// Returns the number of random draws consumed.
__device__ inline void build_row(const DevParams& P, const DevNet& n, const DevVeh& V, const DevLaneDyn& LD, double* rows, int p, int i,
                                 const int* lanes, int nl, int L, double gap, double dn0, double dfree, bool created, double start,
                                 double cpos, double dtv, const int* slm, const MrgA& ma);

// rebuilds the memo of vehicle i (k_ctx_fill)
__device__ inline void ctx_cache_fill(const DevNet& n, const DevVeh& V, int i) {
  CtxCache c;
  const int lane = V.lane[i], route = V.route[i], rpos = V.route_pos[i];
  c.lane = lane; c.id = V.id[i]; c.bits = 0; c.slu0 = -DBL_MAX;
  int cur = lane;
  for (int h = 0; h < 4; h++) {
    int nx = -1, tl = -1;
    if (cur >= 0) nx = next_voie(n, route, rpos, cur, &tl);
    c.chain[h] = nx; c.tl[h] = tl;
    if (nx >= 0 && n.link_brick[n.lane_link[nx]] >= 0) c.bits |= 1 << (8 + h);
    cur = nx;
  }
  for (int k = 0; k < 3; k++) {
    const int ln = k == 0 ? lane : c.chain[k - 1];
    c.slm[k] = -1;
    if (ln < 0) continue;
    const int am = n.link_brick[n.lane_link[ln]] >= 0 ? n.lane_prev[ln] : ln;
    int aval = -1; bool unknown = false;               // first lane after k that is not junction-internal (vehicule.cpp:1854-1874)
    { int q = k; for (; q < 4; q++) { if (c.chain[q] < 0) break; if (!((c.bits >> (8 + q)) & 1)) { aval = c.chain[q]; break; } } unknown = q == 4; }
    int nmatch = 0;
    for (int s = n.sl_off[ln]; s < n.sl_off[ln + 1]; s++) {
      if (!(am >= 0 && n.sl_i[6 * s + 1] == n.lane_link[am] && n.sl_i[6 * s + 2] == n.lane_num[am])) continue;
      if (k == 0) c.slu0 = fmax(c.slu0, n.sl_pos[s]); else if (n.sl_pos[s] >= 0.0) c.bits |= 1 << k;
      if (unknown) { c.slm[k] = -2; continue; }
      if (aval >= 0 && n.sl_i[6 * s + 3] == n.lane_link[aval] && n.sl_i[6 * s + 4] == n.lane_num[aval]) { if (nmatch == 0) c.slm[k] = s; nmatch++; }
    }
    if (nmatch > 1) c.bits |= 1 << (16 + k);
  }
  V.cc[i] = c; V.cc_lane[i] = lane; V.cc_id[i] = c.id;
}
// rows != nullptr: also pack the vehicle's phase-B row at order position p (rows.cuh); the per-vehicle ctx_* arrays are
// then only written for contexts the row cannot hold (general path).
__device__ inline int context_one(const DevParams& P, const DevNet& n, const DevVeh& V, const int* lane_tail, int i,
                                  int own_leader, int* overflow, double* rows = nullptr, int p = -1, const DevLaneDyn* LD = nullptr,
                                  const DevMrg* M = nullptr, double inst = 0.0) {
  const DevCls& c = P.cls[V.cls[i]];
  double p1 = snap_pos(V.pos[i]);
  bool created = p1 <= SYM_UNDEF;                                         // vehicule.cpp:1235
  if (created) p1 = 0.0;                                                  // :1257-1266
  double cpos = created ? V.cpos[i] : 0.0;                                // :1272
  double dtv = P.dt - fmax(0.0, V.tcre[i]);                               // :1278
  double range = max_influence(P, c) + dtv * c.vx;                        // :1710-1715
  int route = V.route[i], rpos = V.route_pos[i];
  int memo[kMemo];
  for (int k = 0; k < kMemo; k++) memo[k] = V.memo[i * kMemo + k];
  int draws = 0;
  int lanes[kMaxCtx]; int nl = 1;
  int lane = V.lane[i];
  lanes[0] = lane;
  double start = p1 + cpos;
  int slm[3] = {-1, -1, -1}; bool fast = false;
  if (rows && V.cc) {                                                     // same hops, same memo touches, read from the vehicle's memo (CtxCache)
    const CtxCache cc = V.cc[i];
    if (cc.lane == lane && cc.id == V.id[i]) {
      int m[kMemo]; for (int k = 0; k < kMemo; k++) m[k] = memo[k];
      int d = 0, h = 0; bool ok = true;
      double dist = len_ex(P, n, lane) - start;
      while (dist < range) {
        if (h > 3) { ok = false; break; }
        const int nx = cc.chain[h], tl = cc.tl[h]; h++;
        if (tl >= 0) d += memo_touch(m, tl);
        if (nx < 0) break;
        if (nl == 3) { ok = false; break; }                               // longer contexts take the general path
        lanes[nl++] = nx; dist += len_ex(P, n, nx);
      }
      for (int k = 0; k < nl && ok; k++) {
        const bool need = k == 0 ? cc.slu0 >= start : (cc.bits >> k) & 1;
        if (need) {
          bool found = false;
          for (int q = k + 1; q < nl; q++) if (!((cc.bits >> (8 + q - 1)) & 1)) found = true;
          for (int hh = nl - 1; !found; hh++) {
            if (hh > 3) { ok = false; break; }
            const int av = cc.chain[hh], tl = cc.tl[hh];
            if (tl >= 0) d += memo_touch(m, tl);
            if (!(av >= 0 && ((cc.bits >> (8 + hh)) & 1))) break;
          }
        }
        int s = cc.slm[k];
        if (s == -2) ok = false;
        else if (s >= 0 && !(n.sl_pos[s] >= (k == 0 ? start : 0.0))) { if ((cc.bits >> (16 + k)) & 1) ok = false; else s = -1; }
        slm[k] = s;
      }
      if (ok) { fast = true; draws = d; for (int k = 0; k < kMemo; k++) memo[k] = m[k]; }
      else nl = 1;
    }
  }
  if (!fast) {
  double dist = len_ex(P, n, lane) - start;                               // CarFollowingContext.cpp:118
  while (dist < range) {                                                  // :121-125
    int tl; int nx = next_voie(n, route, rpos, lane, &tl);
    if (tl >= 0) draws += memo_touch(memo, tl);
    if (nx < 0) break;
    if (nl == kMaxCtx) { atomicAdd(overflow, 1); break; }
    lanes[nl++] = nx; dist += len_ex(P, n, nx); lane = nx;
  }
  }
  // leader: SearchLeaders :128-170 (stops at the first vehicle met, accepted or not)
  int leader = -1; double gap = 0;
  if (own_leader >= 0) {
    double pl = snap_pos(V.pos[own_leader]); if (pl <= SYM_UNDEF) pl = 0.0;
    double g = pl - start;
    if (g <= range) { leader = own_leader; gap = g; }
  } else {
    double d = len_ex(P, n, lanes[0]) - start;
    for (int k = 1; k < nl; k++) {
      int t = lane_tail[lanes[k]];
      if (t >= 0) { double g = snap_pos(V.pos[t]) + d; if (g <= range) { leader = t; gap = g; } break; }
      d += len_ex(P, n, lanes[k]);
    }
  }
  // stop-line bookkeeping: CalculTraficFeux resolves the downstream lane of the movement with CalculNextVoie,
  // which may consume the memoised draw of a lane beyond the context (vehicule.cpp:1854-1874)
  for (int k = 0; k < nl && !fast; k++) {
    int ln = lanes[k];
    double s0 = k == 0 ? start : 0.0;
    bool need = false;
    for (int s = n.sl_off[ln]; s < n.sl_off[ln + 1] && !need; s++) {
      if (n.sl_pos[s] >= s0) {
        int am = n.link_brick[n.lane_link[ln]] >= 0 ? n.lane_prev[ln] : ln;
        if (am >= 0 && n.sl_i[6 * s + 1] == n.lane_link[am] && n.sl_i[6 * s + 2] == n.lane_num[am]) need = true;
      }
    }
    if (!need) continue;
    bool found = false;
    for (int q = k + 1; q < nl && !found; q++) if (n.link_brick[n.lane_link[lanes[q]]] < 0) found = true;
    if (!found) {
      int av = lanes[nl - 1];
      do { int tl; av = next_voie(n, route, rpos, av, &tl); if (tl >= 0) draws += memo_touch(memo, tl); }
      while (av >= 0 && n.link_brick[n.lane_link[av]] >= 0);
    }
  }
  // deltaN at the start of the step (Newell only)
  double dn0 = 1.0;
  if (c.law == 0) {
    int fl = V.flags[i];
    if (leader < 0) dn0 = 1.0;
    else if (!(fl & F_HASCTX)) dn0 = 1.0;
    else if (V.prev_leader[i] < 0 || V.prev_leader[i] != V.id[leader] || (fl & F_PREVFREE)) {
      const DevCls& cl = P.cls[V.cls[leader]];
      double vl1 = V.vit[leader];
      double kv = -cl.kx * cl.w / (vl1 - cl.w);
      dn0 = fmin(1.0, kv * gap);
    } else dn0 = V.dn[i];
  }
  // free-flow distance (Newell) — start position is GetPos(1) without the creation offset (:76)
  double dfree = 0;
  if (c.law == 0) {
    double trav = 0, rem = dtv;
    for (int k = 0; k < nl; k++) {
      int ln = lanes[k]; double sp = k == 0 ? p1 : 0.0; double lex = len_ex(P, n, ln);
      double vdes = fmin(c.vx, n.link_vreg[n.lane_link[ln]]);
      if (P.accbornee) vdes = fmin(vdes, V.vit[i] + acc_max(c, V.vit[i]) * dtv);
      double tt = (n.lane_len[ln] - sp) / vdes;
      if (tt >= rem || k == nl - 1) { trav += rem * vdes; break; }
      trav += lex - sp; rem -= tt;
    }
    dfree = trav;
  }
  // merge points: speed imposed on the head of a non-priority branch, pessimistic speed of such a leader (merge.cuh)
  MrgA ma;
  if (M && p >= 0) { draws += mrg_phase_a(P, n, V, M, i, lanes[0], leader, p1, dtv, inst, memo, &ma); mrg_store(M, p, ma); }
  if (draws) for (int k = 0; k < kMemo; k++) V.memo[i * kMemo + k] = memo[k];
  V.ctx_leader[i] = leader;
  if (!rows || nl > 3) {
    V.ctx_nl[i] = nl;
    for (int k = 0; k < nl; k++) V.ctx_lane[i * kMaxCtx + k] = lanes[k];
    V.ctx_gap[i] = gap; V.ctx_dn0[i] = dn0; V.ctx_dfree[i] = dfree;
    V.ctx_flags[i] = created ? 1 : 0;
  }
  if (rows) build_row(P, n, V, *LD, rows, p, i, lanes, nl, leader, gap, dn0, dfree, created, start, cpos, dtv, fast ? slm : nullptr, ma);
  return draws;
}

// Vehicule::CalculDeceleration (vehicule.cpp:1017-1074); deceleration sigma = 0 => the class value
__device__ inline double decel_limit(const DevParams& P, const DevCls& c, bool ctx_free, double v1, double pas, double limite) {
  double g = limite;
  if (ctx_free) {
    double b = c.decel;
    if (b == 0) return g;
    double N = floor(v1 / (b * P.dt));
    double dd = N * v1 * P.dt - ((N + 1) * N / 2) * b * P.dt * P.dt;
    if (v1 == 0) { double tm = sqrt(2.0 * limite / b); double s = b * tm; g = fmin(g, s * pas); }
    else { double ld = limite - dd; double T = ld / v1;
      if (T <= pas) { if (T < 0) T = 0; g = fmax(fmin(g, v1 * pas - b * (pas - T)), 0.0); } }
  }
  return g;
}

// -------------------------------------------------------------------------------------------------
// Phase B: car-following law + network influences + placement for vehicle i, its leader (if any)
// being already computed unless `leader_computed` is false (loop breaking, vehicule.cpp:1201-1216).
// Vehicule::CalculTraficEx (vehicule.cpp:1299-1574), NewellCarFollowing.cpp:53-68,113-136,204-398,
// GippsCarFollowing.cpp:50-122, IDMCarFollowing.cpp:61-211, Vehicule::CheckNetworkInfluence (:1591-1708),
// Vehicule::CalculTraficFeux (:1799-1945).
// `defer` (relaxation sweeps, flow.cuh): the call may be repeated with another leader speed, so nothing it reads is
// overwritten — creation position from `cpos_in`, creation instant kept, new speed returned through v0_out, the
// desired-lane reset left to phase C (O_VDESRESET).  vL_in: the leader's new speed the caller looked up.
__device__ inline void follow_one(const DevParams& P, const DevNet& n, const DevVeh& V, const DevLaneDyn& LD, int i,
                                  bool leader_computed, double inst, double cpos_in, bool defer = false,
                                  const double* vL_in = nullptr, double* v0_out = nullptr, const DevMrg* M = nullptr, int p = -1, int hdr = 0) {
  const DevCls& c = P.cls[V.cls[i]];
  const bool created = V.ctx_flags[i] & 1;
  double p1 = snap_pos(V.pos[i]); if (p1 <= SYM_UNDEF) p1 = 0.0;
  const double v1 = V.vit[i];
  double cpos = created ? cpos_in : 0.0;
  const double dtv = P.dt - fmax(0.0, V.tcre[i]);
  const int nl = V.ctx_nl[i];
  int lanes[kMaxCtx];
  for (int k = 0; k < nl; k++) lanes[k] = V.ctx_lane[i * kMaxCtx + k];
  const int L = V.ctx_leader[i];
  const double gap = V.ctx_gap[i];
  double dn_end = V.ctx_dn0[i];
  bool ctx_free = true;
  double goal;
  double vL1 = 0;
  if (L >= 0) vL1 = V.vit[L];
  if (c.law == 0) {
    goal = V.ctx_dfree[i];
    if (L >= 0) {
      const DevCls& cl = P.cls[V.cls[L]];
      double vL;
      if (leader_computed) vL = vL_in ? *vL_in : __ldcg(&V.vit_n[L]);   // NewellCarFollowing.cpp:213-216 (written by another SM in the same kernel: read through L2)
      else {                                                                           // :217-230
        double vreg_l = n.link_vreg[n.lane_link[V.lane[L]]];                           // GetVitRegTuyAct of the leader
        if (P.accbornee) vL = fmin(fmin(vreg_l, cl.vx), vL1 + acc_max(c, vL1) * P.dt);
        else vL = fmin(vreg_l, cl.vx);
      }
      if (M && p >= 0) { MrgA ma; mrg_load(M, p, hdr, &ma);                                // NewellCarFollowing.cpp:266-298
        if (ma.lead) { double vit = leader_computed ? vL : ma.vunc; vit = (ma.borne < vit) ? ma.borne : vit; if (ma.posl + vit * dtv > ma.lenl) vL = ma.repl; } }
      const double dn0 = V.ctx_dn0[i];
      // ComputeRelaxationEvolution :113-136 (contraction flag is only set by the aggressiveness procedure: off)
      double kv = -cl.kx * cl.w / (vL1 - cl.w);
      double kvfin = -cl.kx * cl.w / (vL - cl.w);
      double r_self = dn0 / (-c.kx * cl.w);
      double res;
      if ((P.dt - r_self) >= -0.0001) res = fmin(1.0, (dn0 / kv + fmin(r_self / dtv * (vL - vL1) + P.relax, vL) * P.dt) * kvfin);
      else res = fmin(1.0, (dn0 / kv + fmin((vL - vL1) + P.relax, vL) * P.dt) * kvfin);
      double dn1 = fmax(dn0, res);
      dn_end = dn1;
      double r = dn0 / (-cl.kx * cl.w);
      double pc;
      if ((P.dt - r) >= -0.0001) {                                                     // :361-371
        pc = gap + vL * P.dt - dn1 / kvfin;
        if (pc <= 0.0001 && dn0 < 1) { dn_end = dn0; pc = gap + vL * P.dt - dn0 / kvfin; }
      } else {                                                                         // :373-395
        double alpha = -cl.w * cl.kx * P.dt / dn0;
        if (dtv < P.dt) pc = gap + vL * P.dt - dn1 / kvfin;
        else { pc = alpha * gap + vL * P.dt - alpha * dn1 / kvfin;
          if (pc <= 0.0001 && dn0 < 1) { dn_end = dn0; pc = alpha * gap + vL * P.dt - alpha * dn0 / kvfin; } }
      }
      if (pc < goal) { ctx_free = false; goal = pc; }                                  // :60-66
    }
  } else {
    double vmaxl = fmin(c.vx, n.link_vreg[n.lane_link[lanes[0]]]);
    if (c.law == 1) {                                                                  // Gipps
      double vf = v1 + 2.5 * acc_max(c, v1) * dtv * (1 - v1 / vmaxl) * pow_half(0.025 + v1 / vmaxl);
      if (L >= 0) {
        const DevCls& cl = P.cls[V.cls[L]];
        double dmax = -c.decel; if (dmax == 0) dmax = -DBL_MAX;
        double dlmax = -cl.decel; if (dlmax == 0) dlmax = -DBL_MAX;
        double fac = dmax * (dtv / 2.0 + 1);
        double lenL = 1 / cl.kx - cl.esp;
        double sq = fac * fac - dmax * (2.0 * (gap - (lenL + c.esp)) - v1 * dtv - vL1 * vL1 / dlmax);
        double vc = sq >= 0 ? fac + sqrt(sq) : 0.0;
        if (vc < vf) { ctx_free = false; vf = vc; }
      }
      goal = vf * dtv;                                                                 // AbstractCarFollowing.cpp:706-708
    } else {                                                                           // IDM
      double relpos, relsp;
      if (L >= 0) { const DevCls& cl = P.cls[V.cls[L]]; relpos = gap - (1 / cl.kx - cl.esp); relsp = v1 - vL1; }
      else { relpos = DBL_MAX; relsp = v1; }
      double amax = acc_max(c, v1); double dec = c.decel; if (dec == 0) dec = DBL_MAX;
      double z;
      if (relpos <= 0) z = 2;
      else { double dsp = c.esp + fmax(1 * v1 + (v1 * relsp) / (2.0 * sqrt(amax * dec)), 0.0); z = dsp / relpos; }
      ctx_free = z < 1;
      double a = amax * (1.0 - pow_four(v1 / vmaxl) - pow_three_halves(z));
      goal = v1 * dtv + 0.5 * a * dtv * dtv;                                           // :700-703
    }
  }
  bool fluide = ctx_free, feu = false;
  // ---- network influences: vehicule.cpp:1317-1343 -------------------------------------------------
  const double start = p1 + cpos;
  double mgoal = goal + 1, trav = 0;
  const int route = V.route[i], rpos = V.route_pos[i];
  for (int k = 0; k < nl; k++) {
    int ln = lanes[k]; double sp = k == 0 ? start : 0.0; double llen = len_ex(P, n, ln); double ep = goal - trav + sp;
    int link = n.lane_link[ln]; int ta = n.link_type[link];
    double mini = DBL_MAX; bool infl = false;
    if (ep > llen) {                                                                   // 1 - lane end :1602-1616
      bool stop = (n.lane_flags[ln] & 1) || ((k == nl - 1) && ta != 'S' && ta != 'P');
      if (stop) { fluide = false; infl = true; mini = fmin(mini, trav + llen - sp); }
    }
    if (ta == 'S' || ta == 'P') {                                                      // exit capacity :1636-1666
      double pcrit = llen - max_influence(P, c);
      if (ep > pcrit) {
        double ts = LD.next_sortie[ln];
        if (ts < 0) { infl = true; mini = fmin(mini, trav + pcrit - sp); }
        else if (ts > inst - dtv) { double vv = (trav + llen - sp) / (ts - (inst - dtv)); infl = true; mini = fmin(mini, vv * dtv); }
      }
    }
    if (k == 0 && M && p >= 0) { MrgA ma; mrg_load(M, p, hdr, &ma);                     // 3 - Convergent::ComputeTraffic convergent.cpp:1368-1403
      if (ma.own && ma.mc < goal) { infl = true; mini = fmin(mini, ma.mc); } }
    // 4 - signals :1680-1705, CalculTraficFeux :1799-1945
    {
      double ml = goal + 1; bool arret = false, has = false; int aval = -2;
      for (int s = n.sl_off[ln]; s < n.sl_off[ln + 1]; s++) {
        double pf = n.sl_pos[s];
        if (pf >= sp) {
          int am = n.link_brick[link] >= 0 ? n.lane_prev[ln] : ln;
          if (am >= 0 && n.sl_i[6 * s + 1] == n.lane_link[am] && n.sl_i[6 * s + 2] == n.lane_num[am]) {
            if (aval == -2) {
              aval = -1;
              for (int q = k + 1; q < nl && aval < 0; q++) if (n.link_brick[n.lane_link[lanes[q]]] < 0) aval = lanes[q];
              if (aval < 0) { aval = lanes[nl - 1]; int tl;
                do { aval = next_voie(n, route, rpos, aval, &tl); } while (aval >= 0 && n.link_brick[n.lane_link[aval]] >= 0); }
            }
            if (aval >= 0 && n.sl_i[6 * s + 3] == n.lane_link[aval] && n.sl_i[6 * s + 4] == n.lane_num[aval]) {
              int col = n.sl_col[s]; double p = n.sl_prop[s];
              double dloc = (trav - sp) + pf;
              if (col == 0 && fabs(p) < 0.00001) { ml = dloc; arret = true; has = true; break; }
              double tso = v1 > 0 ? dloc / v1 : dloc / (ep / dtv);
              if (col != 1 && p > 0) { if (tso > p * dtv) { ml = dloc; arret = true; has = true; break; } }
              if (col == 1 && p < 1) {
                if (tso > dtv * (1 - p)) {} else { has = true; if (v1 > 0) ml = goal - (1 - p) * v1 * dtv; else ml = goal * p; }
              }
              break;
            }
          }
        }
      }
      if (has) {
        if (arret) { double md = decel_limit(P, c, ctx_free, v1, dtv, ml); ml = fmin(ml, md); }
        if (ml < goal) { infl = true; if (ml < mini) { mini = ml; if (arret) { fluide = false; feu = true; } } }
      }
    }
    if (infl && mini < mgoal) mgoal = mini;
    trav += fmin(ep, llen) - sp;
  }
  goal = fmin(goal, mgoal);                                                            // :1343
  bool cancel = false;
  if (created && goal + cpos <= 0) { cancel = true; cpos = cpos + goal; }              // :1346-1351
  goal = fmax(0.0, goal);
  double v0;
  if (created && !fluide && L >= 0) v0 = fmin(c.vx, vL1);                              // :1358-1361
  else v0 = dtv != 0 ? goal / dtv : c.vx;
  double a0 = created ? 0.0 : (dtv != 0 ? (v0 - v1) / dtv : 0.0);                      // :1368-1375
  if (fabs(a0) < 0.000001) a0 = 0;                                                     // vehicule.h:501
  // ---- placement: vehicule.cpp:1377-1527 ------------------------------------------------------------
  trav = 0; int lane0 = lanes[0]; double pos0 = start; int dest_lane = -1; double trav_dest = 0; bool left = false; int kf = 0;
  const int ro = n.route_off[route], re = n.route_off[route + 1];
  for (int k = 0; k < nl; k++) {
    int ln = lanes[k]; double sp = k == 0 ? start : 0.0; double llen = len_ex(P, n, ln); double ep = goal - trav + sp;
    trav += fmin(ep, llen) - sp;
    int link = n.lane_link[ln]; int ta = n.link_type[link];
    bool is_exit = ta == 'P' || ta == 'S'; bool last = k == nl - 1;
    if ((last && !is_exit && ep > llen) || (feu && ep > llen && ep < llen + 0.00001)) ep = llen;   // :1397-1403
    bool reached = ep > llen && ta == 'S' && n.route_links[re - 1] == link;            // TripLeg.cpp:91-117, Position.cpp:394-404
    if (reached) { dest_lane = ln; trav_dest = trav; }
    if (reached || ep <= llen) { lane0 = ln; pos0 = ep; kf = k; break; }
    left = true;                                                                     // :1505 the desired lane is forgotten when the lane is left
  }
  if (M) mrg_register(n, M, i, p, lanes, kf);                                          // :1485-1502 AddInsVeh, m_LstUsedLanes
  int of = (fluide ? O_FLUIDE : 0) | (feu ? O_FEU : 0) | (ctx_free ? O_CTXFREE : 0) | (created ? O_CREATED : 0);
  if (left) { if (defer) of |= O_VDESRESET; else V.vdes[i] = -1; }
  if (dest_lane >= 0) {                                                                // :1512-1527
    double ti = goal > 0.0001 ? inst - dtv + dtv * (trav_dest / goal) : inst - dtv;
    V.tsort[i] = ti; of |= O_ENDLEG;
  }
  if (cancel) pos0 = SYM_UNDEF;                                                        // :1567-1571
  V.pos_n[i] = pos0; V.acc_n[i] = a0; V.lane_n[i] = lane0;
  if (v0_out) *v0_out = v0; else V.vit_n[i] = v0;
  V.cpos[i] = cpos; if (!defer) V.tcre[i] = 0.0;                                       // :1574
  V.dn[i] = c.law == 0 ? dn_end : 1.0;
  V.prev_leader[i] = L >= 0 ? V.id[L] : -1;
  V.flags[i] = F_HASCTX | (ctx_free ? F_PREVFREE : 0);
  V.oflags[i] = of | (V.status[i] == ST_NEW || V.status[i] == ST_POOL ? O_LINK1NULL : 0);
  (void)ro;
}

}  // namespace symb200
