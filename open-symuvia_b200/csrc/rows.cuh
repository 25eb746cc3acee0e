// rows.cuh — the car-following update split so that the leader-first dependency chain only carries arithmetic.
//
// Vehicule::CalculTraficEx (vehicule.cpp:1192-1588) is one function in the reference; on the device it is
//   phase A  (k_context, parallel)  context, leader, deltaN, free-flow distance AND everything the network
//            influences need that does not depend on the travelled distance: per context lane the length, the
//            "stop at lane end" rule, exit-capacity instant and THE stop line of the vehicle's movement with its
//            colour and green fraction.  Packed into one 256-byte row per vehicle, stored at the vehicle's
//            position in the lane order, so the rows of a segment are contiguous.
//   phase B  (k_relax, flow.cuh: relaxation over the cut leader graph) the leader's new speed -> congested distance -> min
//            with the free-flow distance -> clamps -> travelled distance and speed.  Only reads the row and the leader's
//            current new speed.  NewellCarFollowing.cpp:204-398, vehicule.cpp:1317-1365, :1591-1708, :1799-1945.
//   phase C  (k_place, parallel)  acceleration, placement on the context lanes, destination test, bookkeeping.
//            vehicule.cpp:1367-1527.
// Every formula is evaluated with the same operations in the same order as follow_one() (micro.cuh), which stays
// as the general path for contexts longer than kRowLanes lanes and for the origin queues.
#pragma once
#include "micro.cuh"

namespace symb200 {

constexpr int kRowD = 32;          // doubles per row (256 B)
constexpr int kRowLanes = 3;       // context lanes held inline
// R_T0 = dn0/Kv(v1_L) (Newell) or the law's distance (Gipps/IDM); R_A = -Kx_L*w_L; R_Q1 = (dn0/(-Kx_self*w_L))/dtv: the terms of
// ComputeRelaxationEvolution / ComputeCongestedDistance (NewellCarFollowing.cpp:113-136, :358-395) that do not involve the
// leader's new speed are evaluated in phase A with the same operations, so phase B only keeps the dependent ones.
enum { R_DN0 = 0, R_GAP, R_DFREE, R_DTV, R_V1, R_VL1, R_START, R_CPOS, R_VLEST, R_T0, R_I0, R_I1, R_A, R_WL, R_Q1, R_ALPHA, R_SPARE, R_REC = 17 };
constexpr int R_GOALLAW = R_T0;
constexpr int kRecD = 5;           // llen, pf, prop, ts, ints(lane, flags)
enum { RF_STOPEND = 1, RF_EXIT = 2, RF_SL = 4, RF_COLSHIFT = 4 };
// packed header flags (R_I0 high int)
enum { H_CREATED = 1 << 12, H_FLAGA = 1 << 13, H_FLAGB = 1 << 14, H_SLOW = 1 << 15, H_LAWFREE = 1 << 16, H_LINK1NULL = 1 << 17,
       H_MRGOWN = 1 << 18 /* head of a non-priority merge branch: imposed speed */, H_MRGLEAD = 1 << 19 /* the leader is one */ };
// phase B result flags
enum { B_FLUIDE = 1, B_FEU = 2, B_CTXFREE = 4, B_CANCEL = 8, B_SLOWDONE = 16 };

// Rows live in tiles of 32 consecutive order positions, field-major inside a tile (AoSoA): field k of position p is at
// tile*32*kRowD + k*32 + (p & 31), so a warp reading field k of its 32 vehicles touches one 256-byte line.
struct Row {
  double* b;
  __device__ __forceinline__ double& operator[](int k) const { return b[k * 32]; }
  __device__ __forceinline__ Row operator+(int k) const { return Row{b + k * 32}; }
};
__device__ __forceinline__ Row row_at(const double* rows, int p) { return Row{const_cast<double*>(rows) + (size_t)(p >> 5) * (32 * kRowD) + (p & 31)}; }

__device__ __forceinline__ double pack2(int lo, int hi) { return __hiloint2double(hi, lo); }
__device__ __forceinline__ int lo32(double d) { return __double2loint(d); }
__device__ __forceinline__ int hi32(double d) { return __double2hiint(d); }

// phase A tail: called by k_context once lanes / leader / gap / dn0 / dfree are known for vehicle i at order position p.
__device__ inline void build_row(const DevParams& P, const DevNet& n, const DevVeh& V, const DevLaneDyn& LD, double* rows, int p, int i,
                                 const int* lanes, int nl, int L, double gap, double dn0, double dfree, bool created, double start,
                                 double cpos, double dtv, const int* slm, const MrgA& ma) {
  const Row r = row_at(rows, p);
  const DevCls& c = P.cls[V.cls[i]];
  const double v1 = V.vit[i];
  double vL1 = 0, vlest = 0; int lcls = 0;
  if (L >= 0) {
    vL1 = V.vit[L]; lcls = V.cls[L];
    const DevCls& cl = P.cls[lcls];
    double vreg_l = n.link_vreg[n.lane_link[V.lane[L]]];                 // NewellCarFollowing.cpp:217-230 (leader not computed)
    if (P.accbornee) vlest = fmin(fmin(vreg_l, cl.vx), vL1 + acc_max(c, vL1) * P.dt);
    else vlest = fmin(vreg_l, cl.vx);
  }
  int hdr = V.cls[i] | (lcls << 4) | (nl << 8) | (created ? H_CREATED : 0) | (nl > kRowLanes ? H_SLOW : 0) |
            ((V.status[i] == ST_NEW || V.status[i] == ST_POOL) ? H_LINK1NULL : 0) | (ma.own ? H_MRGOWN : 0) | (ma.lead ? H_MRGLEAD : 0);
  double goal_law = 0;
  if (c.law != 0) {                                                       // Gipps / IDM: no dependency on the leader's new state
    bool ctx_free = true;
    double vmaxl = fmin(c.vx, n.link_vreg[n.lane_link[lanes[0]]]);
    if (c.law == 1) {                                                     // GippsCarFollowing.cpp:50-122
      double vf = v1 + 2.5 * acc_max(c, v1) * dtv * (1 - v1 / vmaxl) * pow_half(0.025 + v1 / vmaxl);
      if (L >= 0) {
        const DevCls& cl = P.cls[lcls];
        double dmax = -c.decel; if (dmax == 0) dmax = -DBL_MAX;
        double dlmax = -cl.decel; if (dlmax == 0) dlmax = -DBL_MAX;
        double fac = dmax * (dtv / 2.0 + 1);
        double lenL = 1 / cl.kx - cl.esp;
        double sq = fac * fac - dmax * (2.0 * (gap - (lenL + c.esp)) - v1 * dtv - vL1 * vL1 / dlmax);
        double vc = sq >= 0 ? fac + sqrt(sq) : 0.0;
        if (vc < vf) { ctx_free = false; vf = vc; }
      }
      goal_law = vf * dtv;
    } else {                                                              // IDMCarFollowing.cpp:61-211
      double relpos, relsp;
      if (L >= 0) { const DevCls& cl = P.cls[lcls]; relpos = gap - (1 / cl.kx - cl.esp); relsp = v1 - vL1; }
      else { relpos = DBL_MAX; relsp = v1; }
      double amax = acc_max(c, v1); double dec = c.decel; if (dec == 0) dec = DBL_MAX;
      double z;
      if (relpos <= 0) z = 2;
      else { double dsp = c.esp + fmax(1 * v1 + (v1 * relsp) / (2.0 * sqrt(amax * dec)), 0.0); z = dsp / relpos; }
      ctx_free = z < 1;
      double a = amax * (1.0 - pow_four(v1 / vmaxl) - pow_three_halves(z));
      goal_law = v1 * dtv + 0.5 * a * dtv * dtv;
    }
    if (ctx_free) hdr |= H_LAWFREE;
  }
  double t0 = goal_law, A = 0, wl = 0, q1 = 0, alpha = 0;
  if (c.law == 0 && L >= 0) {
    const DevCls& cl = P.cls[lcls];
    double kv = -cl.kx * cl.w / (vL1 - cl.w);                              // NewellCarFollowing.cpp:117
    t0 = dn0 / kv; A = -cl.kx * cl.w; wl = cl.w;
    double r_self = dn0 / (-c.kx * cl.w);                                 // :120
    q1 = r_self / dtv;
    if ((P.dt - r_self) >= -0.0001) hdr |= H_FLAGA;                       // :128
    double rr = dn0 / (-cl.kx * cl.w);                                    // :309
    if ((P.dt - rr) >= -0.0001) hdr |= H_FLAGB;                           // :361
    else alpha = -cl.w * cl.kx * P.dt / dn0;                              // :375
  }
  r[R_DN0] = dn0; r[R_GAP] = gap; r[R_DFREE] = dfree; r[R_DTV] = dtv; r[R_V1] = v1; r[R_VL1] = vL1; r[R_START] = start; r[R_CPOS] = cpos;
  // a leader living on another GPU is never computed before its follower: the follower uses the estimated leader speed,
  // exactly as when the reference breaks a leader loop (NewellCarFollowing.cpp:217-230)
  r[R_VLEST] = vlest; r[R_T0] = t0; r[R_I0] = pack2(L, hdr); r[R_I1] = pack2(i, (L >= 0 && V.status[L] == ST_GHOST) ? 1 : 0);
  r[R_A] = A; r[R_WL] = wl; r[R_Q1] = q1; r[R_ALPHA] = alpha;
  if (nl > kRowLanes) return;
  const int route = V.route[i], rpos = V.route_pos[i];
  for (int k = 0; k < nl; k++) {
    int ln = lanes[k]; double sp = k == 0 ? start : 0.0; double llen = len_ex(P, n, ln);
    int link = n.lane_link[ln]; int ta = n.link_type[link];
    int fl = 0; double pf = -1, prop = 0, ts = 0;
    if ((n.lane_flags[ln] & 1) || ((k == nl - 1) && ta != 'S' && ta != 'P')) fl |= RF_STOPEND;        // vehicule.cpp:1607-1608
    if (ta == 'S' || ta == 'P') { fl |= RF_EXIT; ts = LD.next_sortie[ln]; }                          // :1636-1642
    if (slm) {                                                                                        // the movement's stop line from the vehicle's memo (CtxCache)
      const int s = slm[k];
      if (s >= 0) { fl |= RF_SL | (n.sl_col[s] << RF_COLSHIFT); pf = n.sl_pos[s]; prop = n.sl_prop[s]; }
    }
    int aval = -2;                                                                                    // CalculTraficFeux :1829-1883
    for (int s = slm ? n.sl_off[ln + 1] : n.sl_off[ln]; s < n.sl_off[ln + 1]; s++) {
      double f = n.sl_pos[s];
      if (f >= sp) {
        int am = n.link_brick[link] >= 0 ? n.lane_prev[ln] : ln;
        if (am >= 0 && n.sl_i[6 * s + 1] == n.lane_link[am] && n.sl_i[6 * s + 2] == n.lane_num[am]) {
          if (aval == -2) {
            aval = -1;
            for (int q = k + 1; q < nl && aval < 0; q++) if (n.link_brick[n.lane_link[lanes[q]]] < 0) aval = lanes[q];
            if (aval < 0) { aval = lanes[nl - 1]; int tl;
              do { aval = next_voie(n, route, rpos, aval, &tl); } while (aval >= 0 && n.link_brick[n.lane_link[aval]] >= 0); }
          }
          if (aval >= 0 && n.sl_i[6 * s + 3] == n.lane_link[aval] && n.sl_i[6 * s + 4] == n.lane_num[aval]) {
            fl |= RF_SL | (n.sl_col[s] << RF_COLSHIFT); pf = f; prop = n.sl_prop[s];
            break;
          }
        }
      }
    }
    const Row q = r + (R_REC + kRecD * k);
    q[0] = llen; q[1] = pf; q[2] = prop; q[3] = ts; q[4] = pack2(ln, fl);
  }
}

struct BOut { double goal, v0, dn_end, cpos; int flags; };

// phase B: r = the vehicle's row (shared or global memory), vL = leader's new speed (ignored when no leader / non-Newell).
// mx: the merge side record of the position (merge.cuh; read only when the header says so); cutl: the leader stays uncomputed
__device__ __forceinline__ BOut phase_b(const DevParams& P, const Row r, double vL, double inst, const double* mx = nullptr, bool cutl = false) {
  const int hdr = hi32(r[R_I0]); const int L = lo32(r[R_I0]);
  const DevCls& c = P.cls[hdr & 15];
  const int nl = (hdr >> 8) & 15; const bool created = hdr & H_CREATED;
  const double dn0 = r[R_DN0], gap = r[R_GAP], dtv = r[R_DTV], v1 = r[R_V1], vL1 = r[R_VL1], start = r[R_START];
  double cpos = r[R_CPOS];
  double dn_end = dn0; bool ctx_free = true; double goal;
  if (c.law == 0) {
    goal = r[R_DFREE];
    if (L >= 0) {
      if (hdr & H_MRGLEAD) {                                              // NewellCarFollowing.cpp:266-298: the leader heads a non-priority merge branch
        double vit = cutl ? mx[5] : vL; vit = (mx[3] < vit) ? mx[3] : vit;
        if (mx[1] + vit * dtv > mx[2]) vL = mx[4]; }
      const double A = r[R_A];
      double kvfin = A / (vL - r[R_WL]);                                  // NewellCarFollowing.cpp:118, :310
      double res;
      if (hdr & H_FLAGA) res = fmin(1.0, (r[R_T0] + fmin(r[R_Q1] * (vL - vL1) + P.relax, vL) * P.dt) * kvfin);
      else res = fmin(1.0, (r[R_T0] + fmin((vL - vL1) + P.relax, vL) * P.dt) * kvfin);
      double dn1 = fmax(dn0, res);
      dn_end = dn1;
      double pc;
      if (hdr & H_FLAGB) {                                                // :361-371
        pc = gap + vL * P.dt - dn1 / kvfin;
        if (pc <= 0.0001 && dn0 < 1) { dn_end = dn0; pc = gap + vL * P.dt - dn0 / kvfin; }
      } else {                                                            // :373-395
        double alpha = r[R_ALPHA];
        if (dtv < P.dt) pc = gap + vL * P.dt - dn1 / kvfin;
        else { pc = alpha * gap + vL * P.dt - alpha * dn1 / kvfin;
          if (pc <= 0.0001 && dn0 < 1) { dn_end = dn0; pc = alpha * gap + vL * P.dt - alpha * dn0 / kvfin; } }
      }
      if (pc < goal) { ctx_free = false; goal = pc; }
    }
  } else { goal = r[R_GOALLAW]; ctx_free = hdr & H_LAWFREE; dn_end = 1.0; }
  bool fluide = ctx_free, feu = false;
  double mgoal = goal + 1, trav = 0;                                      // vehicule.cpp:1317-1343
  for (int k = 0; k < nl; k++) {
    const Row q = r + (R_REC + kRecD * k);
    const double llen = q[0]; const int fl = hi32(q[4]);
    double sp = k == 0 ? start : 0.0; double ep = goal - trav + sp;
    double mini = DBL_MAX; bool infl = false;
    if (ep > llen && (fl & RF_STOPEND)) { fluide = false; infl = true; mini = fmin(mini, trav + llen - sp); }   // :1602-1616
    if (fl & RF_EXIT) {                                                   // :1636-1666
      double pcrit = llen - P.maxinf[hdr & 15];
      if (ep > pcrit) {
        double ts = q[3];
        if (ts < 0) { infl = true; mini = fmin(mini, trav + pcrit - sp); }
        else if (ts > inst - dtv) { double vv = (trav + llen - sp) / (ts - (inst - dtv)); infl = true; mini = fmin(mini, vv * dtv); }
      }
    }
    if (k == 0 && (hdr & H_MRGOWN)) { const double mc = mx[0]; if (mc < goal) { infl = true; mini = fmin(mini, mc); } }   // Convergent::ComputeTraffic convergent.cpp:1368-1403
    if (fl & RF_SL) {                                                     // :1680-1705, :1884-1938
      double ml = goal + 1; bool arret = false, has = false;
      const int col = (fl >> RF_COLSHIFT) & 3; const double p = q[2]; const double pf = q[1];
      double dloc = (trav - sp) + pf;
      if (col == 0 && fabs(p) < 0.00001) { ml = dloc; arret = true; has = true; }
      else {
        double tso = v1 > 0 ? dloc / v1 : dloc / (ep / dtv);
        if (col != 1 && p > 0 && tso > p * dtv) { ml = dloc; arret = true; has = true; }
        else if (col == 1 && p < 1) {
          if (tso > dtv * (1 - p)) {} else { has = true; if (v1 > 0) ml = goal - (1 - p) * v1 * dtv; else ml = goal * p; }
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
  goal = fmin(goal, mgoal);
  bool cancel = false;
  if (created && goal + cpos <= 0) { cancel = true; cpos = cpos + goal; }  // :1346-1351
  goal = fmax(0.0, goal);
  double v0;
  if (created && !fluide && L >= 0) v0 = fmin(c.vx, vL1);                  // :1358-1361
  else v0 = dtv != 0 ? goal / dtv : c.vx;
  BOut o; o.goal = goal; o.v0 = v0; o.dn_end = dn_end; o.cpos = cpos;
  o.flags = (fluide ? B_FLUIDE : 0) | (feu ? B_FEU : 0) | (ctx_free ? B_CTXFREE : 0) | (cancel ? B_CANCEL : 0);
  return o;
}

// phase C: acceleration, placement, destination, bookkeeping (vehicule.cpp:1367-1527, :1567-1574)
__device__ inline void phase_c(const DevParams& P, const DevNet& n, const DevVeh& V, const double* rows, const double* b_goal,
                               const int* b_flags, int p, double inst, const DevMrg* M = nullptr) {
  const Row r = row_at(rows, p);
  const int bf = b_flags[p];
  const int hdr = hi32(r[R_I0]); const int L = lo32(r[R_I0]); const int i = lo32(r[R_I1]);
  if (bf & B_SLOWDONE) {                                              // placed by follow_one (deferred form): finish its bookkeeping
    V.tcre[i] = 0.0; if (V.oflags[i] & O_VDESRESET) V.vdes[i] = -1;
    return; }
  const int nl = (hdr >> 8) & 15; const bool created = hdr & H_CREATED;
  const double dtv = r[R_DTV], v1 = r[R_V1], start = r[R_START];
  const double goal = b_goal[p];
  const double v0 = V.vit_n[i];
  const bool feu = bf & B_FEU;
  double a0 = created ? 0.0 : (dtv != 0 ? (v0 - v1) / dtv : 0.0);
  if (fabs(a0) < 0.000001) a0 = 0;
  double trav = 0; int lane0 = lo32(r[R_REC + 4]); double pos0 = start; int dest_lane = -1; double trav_dest = 0; int kf = 0;
  const int route = V.route[i]; const int re = n.route_off[route + 1];
  for (int k = 0; k < nl; k++) {
    const Row q = r + (R_REC + kRecD * k);
    int ln = lo32(q[4]); double sp = k == 0 ? start : 0.0; double llen = q[0]; double ep = goal - trav + sp;
    trav += fmin(ep, llen) - sp;
    int link = n.lane_link[ln]; int ta = n.link_type[link];
    bool is_exit = ta == 'P' || ta == 'S'; bool last = k == nl - 1;
    if ((last && !is_exit && ep > llen) || (feu && ep > llen && ep < llen + 0.00001)) ep = llen;
    bool reached = ep > llen && ta == 'S' && n.route_links[re - 1] == link;
    if (reached) { dest_lane = ln; trav_dest = trav; }
    if (reached || ep <= llen) { lane0 = ln; pos0 = ep; kf = k; break; }
    V.vdes[i] = -1;                                                   // vehicule.cpp:1505
  }
  if (M) { int lanes[kRowLanes]; for (int k = 0; k < nl && k < kRowLanes; k++) lanes[k] = lo32(r[R_REC + kRecD * k + 4]);
    mrg_register(n, M, i, p, lanes, kf); }                            // :1485-1502 AddInsVeh, m_LstUsedLanes
  int of = ((bf & B_FLUIDE) ? O_FLUIDE : 0) | (feu ? O_FEU : 0) | ((bf & B_CTXFREE) ? O_CTXFREE : 0) | (created ? O_CREATED : 0);
  if (dest_lane >= 0) {
    double ti = goal > 0.0001 ? inst - dtv + dtv * (trav_dest / goal) : inst - dtv;
    V.tsort[i] = ti; of |= O_ENDLEG;
  }
  if (bf & B_CANCEL) pos0 = SYM_UNDEF;
  V.pos_n[i] = pos0; V.acc_n[i] = a0; V.lane_n[i] = lane0;
  V.tcre[i] = 0.0;
  V.prev_leader[i] = L >= 0 ? V.id[L] : -1;
  V.flags[i] = F_HASCTX | ((bf & B_CTXFREE) ? F_PREVFREE : 0);
  V.oflags[i] = of | ((hdr & H_LINK1NULL) ? O_LINK1NULL : 0);
}

}  // namespace symb200
