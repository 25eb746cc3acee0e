// ======================================================================================
// oracle.cpp — CPU restatement of SymuVia's per-time-step microscopic vehicle update.
//
// *** TEST INFRASTRUCTURE ONLY. ***  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may build, load or call this file.  The product
// (open-symuvia_b200/csrc) never includes, links or executes anything under oracle/.
//
// PARITY UNPINNED: the reference (licit-lab/Open-SymuVia) ships no tests, golden vectors or
// sample networks and cannot be compiled in this image (Boost, Xerces-C,
// XQilla, GDAL, ODBC absent — SURVEY.md §8c).  This restatement follows the cited reference
// code line by line and is pinned only by analytic known-answer tests (tests/test_oracle_kat.py).
// Third-party arithmetic restated from its published algorithm: Boost.Random (version
// unpinned by the reference: conda/meta.yaml:22-23) — mt19937 + uniform_int<>(0,0x7FFE)
// (RandManager.cpp:5,26-30), bucket-rejection of boost/random/uniform_int_distribution.hpp
// (Boost >= 1.47).
//
// Structure mirrors the reference on purpose (AoS vehicles with [0]=end/[1]=start of step,
// per-lane std::set ordered by vehicle id with linear neighbour scans, leader-first
// recursion) so each function can be compared with the file:line it cites.  All paths are
// relative to /root/reference/symuvia/src.
//
// Build: g++ -O2 -ffp-contract=off -std=c++17 -shared -fPIC (see oracle/Makefile).
// ======================================================================================
#include <algorithm>
#include <cassert>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <map>
#include <memory>
#include <set>
#include <string>
#include <vector>

namespace {

constexpr double UNDEF_POSITION = -DBL_MIN;       // symUtil.h:9
constexpr double MAXIMUM_RANDOM_NUMBER = 32767.0; // RandManager.h:9 (0x7fff)
constexpr double NONVAL_DOUBLE = 9999;            // symUtil.h:16

// ---- flat container (format: include/symflat.h) --------------------------------------------
struct Sec { int dtype = 0; std::vector<char> raw; size_t count = 0; };
struct Flat {
  std::map<std::string, Sec> s;
  bool load(const char* path) {
    FILE* f = fopen(path, "rb");
    if (!f) return false;
    char magic[8];
    uint32_t n = 0;
    if (fread(magic, 1, 8, f) != 8 || memcmp(magic, "SYMFLAT1", 8) != 0) { fclose(f); return false; }
    if (fread(&n, 4, 1, f) != 1) { fclose(f); return false; }
    for (uint32_t i = 0; i < n; i++) {
      char name[25] = {0};
      uint32_t dt; uint64_t cnt;
      if (fread(name, 1, 24, f) != 24 || fread(&dt, 4, 1, f) != 1 || fread(&cnt, 8, 1, f) != 1) { fclose(f); return false; }
      size_t isz = dt == 0 ? 4 : dt == 1 ? 8 : 1;
      Sec sec; sec.dtype = (int)dt; sec.count = cnt; sec.raw.resize(cnt * isz);
      if (cnt && fread(sec.raw.data(), 1, sec.raw.size(), f) != sec.raw.size()) { fclose(f); return false; }
      size_t pad = (8 - (sec.raw.size() % 8)) % 8; char tmp[8];
      if (pad && fread(tmp, 1, pad, f) != pad) { fclose(f); return false; }
      s[name] = std::move(sec);
    }
    fclose(f);
    return true;
  }
  const int* I(const char* n) const { auto it = s.find(n); return it == s.end() ? nullptr : (const int*)it->second.raw.data(); }
  const double* D(const char* n) const { auto it = s.find(n); return it == s.end() ? nullptr : (const double*)it->second.raw.data(); }
  size_t N(const char* n) const { auto it = s.find(n); return it == s.end() ? 0 : it->second.count; }
};

// ---- RNG: RandManager.cpp:5-30 --------------------------------------------------------------
struct MT19937 {                                  // boost::mt19937 == std::mt19937 (standard MT19937)
  uint32_t mt[624]; int idx = 624;
  void seed(uint32_t s) { mt[0] = s; for (int i = 1; i < 624; i++) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i; idx = 624; }
  uint32_t next() {
    if (idx >= 624) {
      for (int i = 0; i < 624; i++) {
        uint32_t y = (mt[i] & 0x80000000u) | (mt[(i + 1) % 624] & 0x7fffffffu);
        mt[i] = mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
      }
      idx = 0;
    }
    uint32_t y = mt[idx++];
    y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18;
    return y;
  }
};
struct RandManager {
  MT19937 gen; unsigned count = 0;
  void mySrand(unsigned seed) { gen.seed(seed); count = 0; }       // RandManager.cpp:10-24
  int myRand() {                                                    // RandManager.cpp:26-30
    count++;
    // uniform_int<>(0, 0x7FFE) on a 32-bit engine: range+1 = 32767, bucket = 0xFFFFFFFF / 32767
    // = 131076 (remainder 3 != range so no +1); reject results > range.
    const uint32_t bucket = 0xFFFFFFFFu / 32767u;
    for (;;) { uint32_t r = gen.next() / bucket; if (r <= 32766u) return (int)r; }
  }
};

// ---- static tables ----------------------------------------------------------------------------
struct Cls { double w, kx, vx, esp, decel, tau_lc, pi_rab; int law; std::vector<std::pair<double, double>> plages;
  double debit_max() const { return w * kx / ((w / vx) - 1); }     // DiagFonda.cpp:47
  double k_crit() const { return w * kx / (w - vx); }              // DiagFonda.cpp:48
  double acc_max(double v) const {                                  // vehicule.cpp:157-175
    for (auto& p : plages) if (v <= p.second + 0.000001) return p.first;
    return plages.empty() ? 0.0 : plages.back().first; }
};
struct Link { int first_lane, n_lanes, up_node, down_node; char type_aval; int brick; double length, vit_reg; };
struct Succ { int key_link, next_lane, flags; double rate; };
struct StopLine { int ctrl, up_link, up_num, down_link, down_num; double pos; };
struct Lane { int link, num, prev_lane, flags; double length; std::vector<Succ> succ; std::vector<StopLine> sl;
  double next_inst_sortie = 0; };                                   // voie.h GetNextInstSortie (default 0)
struct Signal { int link_in, link_out; double green, amber, delay; };
struct Seq { double len; std::vector<Signal> sig; };
struct Ctrl { double cycle; std::vector<Seq> seqs; };
struct DestRoute { int route; double coeff; };
struct Dest { int exit; double coeff; std::vector<DestRoute> routes; };
struct Veh;
struct Entry { int node, link; std::vector<std::pair<double, double>> demand; std::vector<Dest> dests;
  std::vector<double> lane_coefs, type_coefs; bool exponential = false;
  double crit = 0, last_inst_creation = 0;                         // m_dbCritCreationVeh, m_dbLastInstCreation
  std::vector<int> pret;                                           // m_mapVehPret[link][lane] (-1 none)
  std::vector<std::deque<Veh*>> attente; };                        // m_mapVehEnAttente[link][lane]
struct Exit { int node, link; double capacity; };
// ---- merge points (row A10): convergent.h, sensors/PonctualSensor.h, sensors/SensorsManager.h ----
struct DesVoieAmont { int lane, niv, niv_ref; };                   // convergent.h DesVoieAmont
struct PtCvg { int cvg, down_lane; std::vector<DesVoieAmont> vam;  // PointDeConvergence
  double inst_last_passage = 0; int rand_numbers = 0; bool free_flow = true, insertion = false; };
struct Capteur { int link; double pos; std::vector<double> cnt; }; // PonctualSensor::m_Data.pNbTypeVehInst [cls][lane][window]
struct Cvg { int node, down_link; double tagreg, gamma, mu, pos_cpt_av; std::vector<double> tf; std::vector<int> pts, am_links;
  std::vector<Capteur> cpt; int nb_inst = -1, last_ind = 0, window = 0;                      // SensorsManager::m_nNbInst (-1 at construction)
  std::vector<Veh*> ins_veh; };                                    // ConnectionPonctuel::m_LstInsVeh
struct CafMvt { int node, ent_link, exit_link; std::vector<int> tl; };   // CarrefourAFeuxEx MvtCAFEx (entry, exit, lstTMvt)

// ---- car-following context: carFollowing/CarFollowingContext.h, NewellContext.h ---------------
struct Ctx {
  std::vector<int> lanes; Veh* leader = nullptr; int leader_id = -1; double leader_dist = 0;
  double start_pos = 0; bool free_flow = true; double deltaN = 1; bool contraction = false;
  bool post = false;                                               // CarFollowingContext::m_bIsPostProcessing
};

// ---- vehicle: vehicule.h:310-420 --------------------------------------------------------------
struct Veh {
  int id = 0, cls = 0; double w = 0, kx = 0, vx = 0, vit_max = 0;
  double pos[2], vit[2], acc[2]; int lane[2], link[2];
  int route = -1, route_pos = 0, next_link = -1;
  bool deja_calcule = false, regime_fluide = true, regime_fluide_leader = true, arret_au_feu = false;
  bool end_leg = false, b_deceleration = false, attente_insertion = false;
  double deceleration = -1;
  double inst_creation = 0, creation_pos = 0, generation_pos = UNDEF_POSITION, heure_entree = 0, inst_sortie = -1;
  double inst_dest_reached = 0; int dest_enter_lane = -1;
  double dst_parcourue = 0, vit_reg_tuy_act = 0;
  int voie_desiree = -1; int prev_lane = -1; bool b_chgt_voie = false; std::vector<char> voies_ok;
  std::vector<int> used_lanes; std::map<int, double> tirage;      // m_LstUsedLanes, m_NextVoieTirage
  std::vector<int> tuyaux_parcourus;
  std::unique_ptr<Ctx> ctx, prev_ctx;
  bool in_list = false; int dead = 0;
  int next_voie = -1;                                              // m_pNextVoie: set by ChangementVoie (vehicule.cpp:2700) and CalculVitesseLeader (convergent.cpp:1146), never reset
  double res_tir_foll_int = -1;                                    // m_dbResTirFollInt (AbstractCarFollowing.cpp:484-497)
};
struct LessId { bool operator()(const Veh* a, const Veh* b) const { return a->id < b->id; } };   // tools.h:67-74
#ifdef ORACLE_FLAT_LANES
// "sorted lanes" variant of the CPU baseline: the per-lane vehicle set is a flat array kept in id order instead of a
// red-black tree.  Same iteration order, hence the same results; what changes is the memory behaviour of InitVehicules and of the scans.
struct LaneSet {
  std::vector<Veh*> v;
  void clear() { v.clear(); }
  void insert(Veh* x) { auto it = std::lower_bound(v.begin(), v.end(), x, LessId()); if (it == v.end() || (*it)->id != x->id) v.insert(it, x); }
  void erase(Veh* x) { auto it = std::lower_bound(v.begin(), v.end(), x, LessId()); if (it != v.end() && *it == x) v.erase(it); }
  std::vector<Veh*>::const_iterator begin() const { return v.begin(); }
  std::vector<Veh*>::const_iterator end() const { return v.end(); }
};
#else
typedef std::set<Veh*, LessId> LaneSet;
#endif

struct Exited { int id; double inst; };

// =================================================================================================
struct Sim {
  // parameters
  double dt = 1, duree = 0; unsigned seed = 0; bool accbornee = false; double relax = 1; bool chgtvoie = false;
  double dst_chgtvoie = 50; double bevel = 0; double dst_chgtvoie_force = -1, phi_chgtvoie_force = 1; int ordre_chgtvoie = 1;
  std::vector<Cls> cls; std::vector<Link> links; std::vector<Lane> lanes; std::vector<Ctrl> ctrls;
  std::vector<std::vector<int>> routes; std::vector<Entry> entries; std::vector<Exit> exits;
  std::vector<int> node_type, node_ctrl;
  std::vector<Cvg> cvgs; std::vector<PtCvg> pts; std::vector<CafMvt> caf_mvts;              // Reseau::Liste_convergents in std::map (id) order
  std::vector<int> link_cvg_aval;                                   // link -> Convergent that is its downstream connection (-1)
  std::vector<std::vector<std::pair<int, int>>> link_capteurs;      // Tuyau::getListeCapteursConvergents: (cvg, sensor)
  std::vector<std::vector<int>> lcam, lcav;                         // m_LstTuyAm of a link's upstream connection / m_LstTuyAv of its downstream one
  std::vector<std::vector<int>> node_mvts;                          // CarrefourAFeuxEx::m_LstMouvements per junction (indices into caf_mvts, reference order)
  long n_insertions = 0, n_refus = 0, n_pins = 0, n_cong = 0; double inst_cur = 0;
  bool trace_ties = getenv("ORACLE_TRACE_TIES") != nullptr, trace_merge = getenv("ORACLE_TRACE_MERGE") != nullptr;
  // dynamic
  double t = 0; int ninst = 0; int last_id = 0; RandManager rnd;
  std::vector<Veh*> lst;                                           // m_LstVehicles (order is part of the contract)
  std::vector<std::unique_ptr<Veh>> owned;                         // everything ever created and still alive
  std::vector<LaneSet> lane_set;                                   // VoieMicro::m_LstVehiculesAnt (voie.h:240)
  std::vector<int> nb_entre, nb_sorti;                            // Tuyau::m_mapNbVehEntre/Sorti (tuyau.h:531-532), [link*T+cls]
  std::vector<Exited> exited;                                      // removed this step
  std::vector<Veh*> graveyard;
  double min_kv = 0, max_vit_max = 0, max_debit_max = 0;
  long n_ties = 0, n_loops = 0, n_created = 0, max_depth = 0;
  // partitioned model (multi-GPU runs): a leader on a lane of another partition is never computed first; its follower uses the
  // estimated leader speed of NewellCarFollowing.cpp:217-230 (the reference's own rule when a leader is not yet computed)
  std::vector<int> lane_owner;
  bool same_part(const Veh* a, const Veh* b) const { return lane_owner.empty() || lane_owner[a->lane[1]] == lane_owner[b->lane[1]]; }

  struct ApiVeh { int id, cls, entry, dest, lane; double dbt; };
  std::vector<ApiVeh> api_queue;                                   // Reseau::m_VehiclesToCreate (reseau.cpp:2233-2240)

  // ---------------------------------------------------------------------------------------------
  bool load(const char* path) {
    Flat f; if (!f.load(path)) return false;
    const double* P = f.D("params");
    dt = P[0]; duree = P[1] * dt; seed = (unsigned)P[2]; accbornee = P[3] != 0; relax = P[4]; chgtvoie = P[5] != 0;
    dst_chgtvoie = P[6]; phi_chgtvoie_force = P[9] > 0 ? P[9] : 1.0; dst_chgtvoie_force = P[10] > 0 ? P[10] : -1.0;   // reseau.cpp:272-273, :5601-5604
    const double* C = f.D("cls"); size_t T = f.N("cls") / 24;
    for (size_t i = 0; i < T; i++) { const double* c = C + 24 * i; Cls k; k.w = c[0]; k.kx = c[1]; k.vx = c[2]; k.esp = c[3];
      k.decel = c[4]; k.tau_lc = c[5]; k.pi_rab = c[6]; k.law = (int)c[7];
      for (int j = 0; j < (int)c[8]; j++) k.plages.push_back({c[9 + 2 * j], c[10 + 2 * j]});
      cls.push_back(k); }
    const int* li = f.I("link_i"); const double* lf = f.D("link_f"); size_t K = f.N("link_i") / 8;
    for (size_t k = 0; k < K; k++) links.push_back({li[8 * k], li[8 * k + 1], li[8 * k + 2], li[8 * k + 3], (char)li[8 * k + 4], li[8 * k + 5], lf[2 * k], lf[2 * k + 1]});
    const int* la = f.I("lane_i"); const double* laf = f.D("lane_f"); size_t L = f.N("lane_i") / 4;
    const int* so = f.I("succ_off"); const int* si = f.I("succ_i"); const double* sf = f.D("succ_f");
    const int* lo = f.I("sl_off"); const int* sli = f.I("sl_i"); const double* slf = f.D("sl_f");
    for (size_t l = 0; l < L; l++) { Lane x; x.link = la[4 * l]; x.num = la[4 * l + 1]; x.prev_lane = la[4 * l + 2]; x.flags = la[4 * l + 3]; x.length = laf[l];
      for (int m = so[l]; m < so[l + 1]; m++) x.succ.push_back({si[3 * m], si[3 * m + 1], si[3 * m + 2], sf[m]});
      for (int m = lo[l]; m < lo[l + 1]; m++) x.sl.push_back({sli[6 * m], sli[6 * m + 1], sli[6 * m + 2], sli[6 * m + 3], sli[6 * m + 4], slf[m]});
      lanes.push_back(std::move(x)); }
    const int* ci = f.I("ctrl_i"); const double* cf = f.D("ctrl_f"); const int* qi = f.I("seq_i"); const double* qf = f.D("seq_f");
    const int* gi = f.I("sig_i"); const double* gf = f.D("sig_f");
    for (size_t c = 0; c < f.N("ctrl_i") / 2; c++) { Ctrl x; x.cycle = cf[c];
      for (int q = ci[2 * c]; q < ci[2 * c] + ci[2 * c + 1]; q++) { Seq s; s.len = qf[q];
        for (int g = qi[2 * q]; g < qi[2 * q] + qi[2 * q + 1]; g++) s.sig.push_back({gi[2 * g], gi[2 * g + 1], gf[3 * g], gf[3 * g + 1], gf[3 * g + 2]});
        x.seqs.push_back(std::move(s)); }
      ctrls.push_back(std::move(x)); }
    const int* ni = f.I("node_i");
    for (size_t n = 0; n < f.N("node_i") / 4; n++) { node_type.push_back(ni[4 * n]); node_ctrl.push_back(ni[4 * n + 1]); }
    const int* ro = f.I("route_off"); const int* rl = f.I("route_links");
    for (size_t r = 0; r + 1 < f.N("route_off"); r++) routes.emplace_back(rl + ro[r], rl + ro[r + 1]);
    const int* xi = f.I("exit_i"); const double* xf = f.D("exit_f");
    for (size_t x = 0; x < f.N("exit_i") / 2; x++) exits.push_back({xi[2 * x], xi[2 * x + 1], xf[x]});
    const int* ei = f.I("entry_i"); const double* df = f.D("dem_f"); const int* di = f.I("dest_i"); const double* dcf = f.D("dest_f");
    const int* dri = f.I("droute_i"); const double* drf = f.D("droute_f"); const double* lcf = f.D("lanecoef_f"); const double* tcf = f.D("typecoef_f");
    for (size_t e = 0; e < f.N("entry_i") / 8; e++) { const int* r = ei + 8 * e; Entry x; x.node = r[0]; x.link = r[1];
      for (int d = r[2]; d < r[2] + r[3]; d++) x.demand.push_back({df[2 * d], df[2 * d + 1]});
      for (int d = r[4]; d < r[4] + r[5]; d++) { Dest y; y.exit = di[3 * d]; y.coeff = dcf[d];
        for (int q = di[3 * d + 1]; q < di[3 * d + 1] + di[3 * d + 2]; q++) y.routes.push_back({dri[q], drf[q]});
        x.dests.push_back(std::move(y)); }
      int nl = links[x.link].n_lanes; x.exponential = (r[7] >> 30) & 1; int tco = r[7] & 0x3fffffff;
      for (int k = 0; k < nl; k++) x.lane_coefs.push_back(lcf[r[6] + k]);
      for (size_t k = 0; k < T; k++) x.type_coefs.push_back(tcf[tco + k]);
      x.pret.assign(nl, -1); x.attente.resize(nl);
      entries.push_back(std::move(x)); }
    lane_set.resize(lanes.size()); nb_entre.assign(links.size() * T, 0); nb_sorti.assign(links.size() * T, 0);
    // merge points: Convergent::Init / FinInit (convergent.cpp:208-255), CarrefourAFeuxEx::Init (:489-919, :1129-1166)
    link_cvg_aval.assign(links.size(), -1); link_capteurs.resize(links.size()); lcam.resize(links.size()); lcav.resize(links.size());
    if (f.N("cvg_i")) {
      const int* ci2 = f.I("cvg_i"); const double* cf2 = f.D("cvg_f"); const double* ctf = f.D("cvg_tf"); const int* cam = f.I("cvgam_i");
      const int* pi = f.I("pt_i"); const int* pam = f.I("ptam_i");
      for (size_t p = 0; p < f.N("pt_i") / 4; p++) { PtCvg x; x.cvg = pi[4 * p]; x.down_lane = pi[4 * p + 1];
        for (int m = pi[4 * p + 2]; m < pi[4 * p + 2] + pi[4 * p + 3]; m++) x.vam.push_back({pam[2 * m], pam[2 * m + 1], pam[2 * m + 1]});
        pts.push_back(std::move(x)); }
      for (size_t c = 0; c < f.N("cvg_i") / 8; c++) { const int* r = ci2 + 8 * c; Cvg x; x.node = r[0]; x.down_link = r[1];
        x.tagreg = cf2[4 * c]; x.gamma = cf2[4 * c + 1]; x.mu = cf2[4 * c + 2]; x.pos_cpt_av = cf2[4 * c + 3];
        for (size_t t = 0; t < T; t++) x.tf.push_back(ctf[c * T + t]);
        for (int p = r[2]; p < r[2] + r[3]; p++) x.pts.push_back(p);
        for (int a = r[4]; a < r[4] + r[5]; a++) x.am_links.push_back(cam[a]);
        x.window = (int)(x.tagreg / dt);                                                       // PonctualSensor::InitData :131
        auto add = [&](int link, double pos) { Capteur k; k.link = link; k.pos = pos; k.cnt.assign(T * links[link].n_lanes * x.window, 0.0);
          link_capteurs[link].push_back({(int)c, (int)x.cpt.size()}); x.cpt.push_back(std::move(k)); };
        add(x.down_link, x.pos_cpt_av);                                                        // "CptTav"
        for (int a : x.am_links) { add(a, links[a].length - 5); link_cvg_aval[a] = (int)c; }   // "CptAm" at length - 5
        cvgs.push_back(std::move(x)); }
      const int* mi = f.I("cafmvt_i"); const int* ml = f.I("cafmvt_links");
      for (size_t m = 0; m < f.N("cafmvt_i") / 5; m++) { CafMvt x{mi[5 * m], mi[5 * m + 1], mi[5 * m + 2], {}};
        for (int q = mi[5 * m + 3]; q < mi[5 * m + 3] + mi[5 * m + 4]; q++) x.tl.push_back(ml[q]);
        caf_mvts.push_back(std::move(x)); }
      node_mvts.resize(node_type.size());
      for (size_t m = 0; m < caf_mvts.size(); m++) node_mvts[caf_mvts[m].node].push_back((int)m);
    }
    if (f.N("lcam_off")) { const int* o = f.I("lcam_off"); const int* l = f.I("lcam_links"); for (size_t k = 0; k < links.size(); k++) lcam[k].assign(l + o[k], l + o[k + 1]); }
    if (f.N("lcav_off")) { const int* o = f.I("lcav_off"); const int* l = f.I("lcav_links"); for (size_t k = 0; k < links.size(); k++) lcav[k].assign(l + o[k], l + o[k + 1]); }
    // Reseau::GetMinKv reseau.cpp:9719-9737, GetMaxVitMax :9700-9715, m_dbMaxDebitMax :5915
    min_kv = 999; max_vit_max = 0; max_debit_max = 0;
    for (auto& c : cls) { double kv = -c.kx * c.w / (c.vx - c.w); if (kv < min_kv) min_kv = kv; if (c.vx > max_vit_max) max_vit_max = c.vx;
      if (c.debit_max() > max_debit_max) max_debit_max = c.debit_max(); }
    rnd.mySrand(seed);                                              // reseau.cpp:983-995 (seed != 0)
    // initial vehicles: reseau.cpp:11680-11773 (load), :1293-1340 (copy into m_LstVehicles at InitSimuTrafic)
    const int* ivi = f.I("iv_i"); const double* ivf = f.D("iv_f"); const int* ivid = f.N("iv_id") ? f.I("iv_id") : nullptr;
    for (size_t v = 0; v < f.N("iv_i") / 4; v++) {
      if (ivid) last_id = ivid[v];
      Veh* p = new_vehicle(ivi[4 * v], last_id++, false);   // law object copied, no draw (vehicule.cpp:5246-5250)
      p->lane[0] = ivi[4 * v + 1]; p->link[0] = lanes[p->lane[0]].link; p->route = ivi[4 * v + 2]; p->route_pos = ivi[4 * v + 3];
      p->pos[0] = ivf[3 * v]; p->vit[0] = ivf[3 * v + 1]; p->acc[0] = ivf[3 * v + 2];
      p->vit_reg_tuy_act = links[p->link[0]].vit_reg; p->heure_entree = 0;
      add_vehicule(p);
    }
    return true;
  }

  // Vehicule::Vehicule vehicule.cpp:417-483 (state = UNDEF / vitmax / 0), law pick CarFollowingFactory.cpp:88-137
  Veh* new_vehicle(int c, int id, bool draw_law = true) {
    auto v = std::make_unique<Veh>(); Veh* p = v.get();
    p->id = id; p->cls = c; p->w = cls[c].w; p->kx = cls[c].kx; p->vx = cls[c].vx; p->vit_max = cls[c].vx;
    for (int i = 0; i < 2; i++) { p->pos[i] = UNDEF_POSITION; p->vit[i] = p->vit_max; p->acc[i] = 0; p->lane[i] = -1; p->link[i] = -1; }
    if (draw_law && cls[c].law != 0) rnd.myRand();       // scripted-law coefficient pick draws once: CarFollowingFactory.cpp:99
    owned.push_back(std::move(v));
    return p;
  }
  void add_vehicule(Veh* p) { lst.push_back(p); p->in_list = true; }   // Reseau::AddVehicule reseau.cpp:4386-4395

  // ------------------------------------------------------------------------------------------
  // small accessors
  double len_ex(int l) const { return (lanes[l].flags & 1) ? lanes[l].length - bevel : lanes[l].length; }   // voie.cpp:1451-1457
  bool chgt_voie_obligatoire(int l) const { return lanes[l].flags & 1; }                                   // voie.cpp:789-810
  double vit_reg(int link) const { return links[link].vit_reg; }           // tuyau.cpp:951-977 (no per-portion exceptions in scope)
  double max_influence(const Veh* v) const {
    const Cls& c = cls[v->cls];
    if (c.law == 0) return 1.0 / min_kv;                                   // NewellCarFollowing.cpp:48-51
    if (c.law == 1) return 3 * v->vit_max * dt;                           // GippsCarFollowing.cpp:45-48
    double dec = c.decel; if (dec == 0) dec = DBL_MAX;                     // IDMCarFollowing.cpp:49-59
    double s = c.esp + std::max<double>(1 * v->vit_max + v->vit_max * v->vit_max / (2.0 * sqrt(c.acc_max(v->vit_max) * dec)), 0);
    return 3 * s;
  }
  double veh_length(const Veh* v) const { return 1 / v->kx - cls[v->cls].esp; }   // AbstractCarFollowing.cpp:50-53

  // ------------------------------------------------------------------------------------------
  // signals: ControleurDeFeux.cpp:203-307, :1517-1558, :96-148
  int couleur(const Ctrl& c, double inst, int te, int ts, double* prop) const {   // 0 red 1 green 2 amber
    int nsec = (int)inst;                                   // STimeSpan((int)dbInstant), plan starts with the simulation
    int ntmp = nsec % (int)c.cycle;                         // :250-251
    double prev = 0; const Seq* seq = nullptr; double tcur = 0;
    for (auto& s : c.seqs) { if (prev + s.len > ntmp) { tcur = ntmp - prev; seq = &s; break; } prev += s.len; }   // :1209-1228
    if (!seq) return 1;                                     // :280-281 (no sequence -> green)
    const Signal* sg = nullptr;
    for (auto& g : seq->sig) if (g.link_in == te && g.link_out == ts) { sg = &g; break; }
    if (!sg) return 0;                                      // :293-295
    int col;
    if (tcur < sg->delay) col = 0; else if (tcur < sg->delay + sg->green) col = 1;
    else if (sg->delay + sg->green <= tcur && tcur < sg->delay + sg->green + sg->amber) col = 2; else col = 0;   // :1517-1541
    if (prop) {
      if (col != 0) *prop = std::min<double>(1, (tcur - sg->delay) / dt);                                      // :1544-1549
      else if (tcur < sg->delay) *prop = 1 - std::min<double>(1, tcur / dt);
      else *prop = 1 - std::min<double>(1, (tcur - (sg->delay + sg->green + sg->amber) / dt));                 // :1555 (sic)
    }
    return col;
  }
  int statut_couleur(int ctrl, double inst, int te, int ts, double* prop) const {     // GetStatutCouleurFeux :96-148
    const Ctrl& c = ctrls[ctrl];
    int col = couleur(c, inst, te, ts, prop);
    int pcol = (inst - dt > 0) ? couleur(c, inst - dt, te, ts, nullptr) : 1;
    if (col == 0 && pcol == 0) { *prop = 0; return 0; }
    if ((col == 1 || col == 2) && (pcol == 1 || pcol == 2)) { *prop = 1; return col; }
    if (col == 1 && pcol == 0) return 1;
    if (col == 0 && (pcol == 1 || pcol == 2)) { *prop = 1; return 0; }
    return 0;
  }

  // ------------------------------------------------------------------------------------------
  // routing: Vehicule::CalculNextTuyau vehicule.cpp:2764-2826 ('iti' mode, predefined path)
  // Vehicule::GetItineraireIndex(isSameTuyau): first index of `link` in the route.  Routes in scope never
  // repeat a link, so searching forward from the cursor first returns the same index.
  int route_index(const Veh* v, int link) const {
    const auto& r = routes[v->route];
    for (size_t k = (size_t)std::max(0, v->route_pos); k < r.size(); k++) if (r[k] == link) return (int)k;
    for (size_t k = 0; k < r.size() && (int)k < v->route_pos; k++) if (r[k] == link) return (int)k;
    return -1;
  }
  int calcul_next_tuyau(Veh* v, int link) {
    if (v->next_link >= 0 && link >= 0 && links[link].brick >= 0) return v->next_link;       // :2784-2789
    const auto& r = routes[v->route];
    // internal link: use the route link upstream of the brick (:2798-2806) = link of the movement's entry lane
    if (links[link].brick >= 0) link = lanes[lanes[links[link].first_lane].prev_lane].link;
    int i = route_index(v, link);                                                            // :2808
    if (i < 0) return -1;
    return (i + 1 < (int)r.size()) ? r[i + 1] : -1;                                          // :2811-2823
  }
  // Vehicule::GetTirage vehicule.cpp:1779-1795
  double get_tirage(Veh* v, int lane_am) {
    auto it = v->tirage.find(lane_am);
    if (it != v->tirage.end()) return it->second;
    double r = rnd.myRand() / MAXIMUM_RANDOM_NUMBER;
    v->tirage[lane_am] = r; return r;
  }
  // Vehicule::CalculNextVoie vehicule.cpp:3070-3469, 'iti' branch (:3347-3366, :3384-3461) on flattened
  // movements: succ rows of a lane = allowed (next external link -> lane) with rates (Connection.h
  // GetCoeffsMouvementAutorise); for an internal CAF lane the key is the exit link and the memo key is the
  // entry lane (pVAm, :3389), CarrefourAFeuxEx::GetNextTroncon resolves to the single successor.
  int calcul_next_voie(Veh* v, int lane) {
    if (lane < 0) return -1;
    if (chgt_voie_obligatoire(lane)) return -1;                                               // :3105-3108
    const Lane& L = lanes[lane]; const Link& T = links[L.link];
    if (T.type_aval == 'S' || T.type_aval == 'P' || T.type_aval == 'E') return -1;            // no pCnx -> NULL (:3463)
    const auto& r = routes[v->route];
    int ref_link = L.link; int lane_am = lane;
    if (T.brick >= 0) { lane_am = L.prev_lane; ref_link = lanes[lane_am].link; }              // :3172-3199 (GetVoieAmont)
    int i = route_index(v, ref_link);                                                         // GetItineraireIndex(isSameTuyau) :3350
    if (i < 0 || i + 1 >= (int)r.size()) return -1;                                           // :3350-3357
    int ts = r[i + 1];
    // allowed movement lookup
    double sum = 0; bool any = false;
    for (auto& s : L.succ) if (s.key_link == ts) { any = true; break; }
    if (!any) return -1;                                                                      // bMvtAutorise false (:3446-3449)
    double rand = get_tirage(v, lane_am);                                                     // :3413
    for (auto& s : L.succ) if (s.key_link == ts && s.rate > 0) { sum += s.rate; if (rand <= sum) return s.next_lane; }   // :3421-3444
    return -1;
  }

  // ------------------------------------------------------------------------------------------
  // leader searches: reseau.cpp:3583-3640 (GetNearAvalVehicule), :3650-3820 (GetPrevVehicule),
  // tie-breakers :14041-14128
  Veh* get_leader_of(Veh* v) const { return (v->ctx && v->ctx->leader) ? v->ctx->leader : nullptr; }   // vehicule.cpp:5636-5644
  Veh* get_last_vehicule(std::vector<Veh*>& l) {                                            // :14041-14081
    if (l.size() == 1) return l[0];
    if (l.empty()) return nullptr;
    n_ties++;
    if (trace_ties) { fprintf(stderr, "tie t=%g lane %d (link %d brick %d):", t, l[0]->lane[1], lanes[l[0]->lane[1]].link, links[lanes[l[0]->lane[1]].link].brick); for (Veh* q : l) fprintf(stderr, " id %d pos %.6f v1 %.4f", q->id, q->pos[1], q->vit[1]); fprintf(stderr, " len %.4f\n", lanes[l[0]->lane[1]].length); }
    for (size_t i = 0; i < l.size(); i++) { bool fol = false;
      for (size_t j = 0; j < l.size() && !fol; j++) if (i != j) { Veh* ld = get_leader_of(l[j]); if (ld && ld->id == l[i]->id) fol = true; }
      if (!fol) return l[i]; }
    return l.back();
  }
  Veh* get_near_aval_vehicule(int lane, double pos, Veh* excl) {                            // :3583-3640
    std::vector<Veh*> l; if (lane < 0) return nullptr;
    double pv = lanes[lane].length + 0.00001;
    for (Veh* c : lane_set[lane]) {
      if (c->pos[1] >= pos && c->pos[1] <= pv) { if (!excl || excl->id != c->id) {
        if (c->pos[1] == pv) l.push_back(c); else { l.resize(1); l[0] = c; pv = c->pos[1]; } } } }
    return get_last_vehicule(l);
  }
  Veh* get_prev_vehicule(Veh* v) {                                                          // :3650-3718 (bSearchNextVoie=false)
    int lane = v->lane[1]; if (lane < 0) return nullptr;
    std::vector<Veh*> l; double pv = lanes[lane].length + 1;
    for (Veh* c : lane_set[lane]) {
      if ((c->pos[1] > v->pos[1] || (v->pos[0] == UNDEF_POSITION && c->pos[1] == v->pos[1])) && c->pos[1] <= pv && c != v) {
        if (c->pos[1] == pv) l.push_back(c); else { l.resize(1); l[0] = c; pv = c->pos[1]; } } }
    return get_last_vehicule(l);
  }


  // ------------------------------------------------------------------------------------------
  // lane changing (row A9): reseau.cpp:4219-4330, vehicule.cpp:2217-2456, :2657-2711, :4749-5039,
  // carFollowing/AbstractCarFollowing.cpp:61-329, reseau.cpp:3255-3295, :3374-3520, :4446-4560, :14084-14128
  bool mvt_autorise(int lane, int next_link) const { if (next_link < 0) return false; for (auto& q : lanes[lane].succ) if (q.key_link == next_link) return true; return false; }
  void calcul_voies_possibles(Veh* v) {                                                      // vehicule.cpp:4749-4940
    if (v->link[1] < 0 || v->lane[1] < 0) return;
    const Link& T = links[v->link[1]]; v->voies_ok.assign(T.n_lanes, 0);
    auto plain = [&]() { for (int i = 0; i < T.n_lanes; i++) v->voies_ok[i] = !chgt_voie_obligatoire(T.first_lane + i); };
    if (v->pos[1] > lanes[v->lane[1]].length - dst_chgtvoie) {                               // :4805 mandatory-changing zone
      bool has_cnx = T.type_aval != 0 && T.brick < 0;                                        // :4809-4830 (CAF downstream, or a simple connection)
      if (has_cnx) {
        if (T.type_aval == 'S' || T.type_aval == 'P') { for (int i = 0; i < T.n_lanes; i++) v->voies_ok[i] = 1; return; }   // :4835-4846
        for (int i = 0; i < T.n_lanes; i++) v->voies_ok[i] = mvt_autorise(T.first_lane + i, v->next_link) && !chgt_voie_obligatoire(T.first_lane + i);   // :4892-4909
      } else plain();
    } else plain();
  }
  void calcul_voie_desiree(Veh* v) {                                                         // vehicule.cpp:4942-5038
    if (v->voie_desiree != -1) return;
    if (v->lane[1] < 0) { v->voie_desiree = -1; return; }
    if (v->voies_ok.empty()) return;
    int nV = lanes[v->lane[1]].num; int n = (int)v->voies_ok.size();
    if (!v->voies_ok[nV]) {
      if (nV - 1 >= 0 && v->voies_ok[nV - 1]) { v->voie_desiree = nV - 1; return; }
      if (nV + 1 < n && v->voies_ok[nV + 1]) { v->voie_desiree = nV + 1; return; }
      for (int k = nV - 2; k >= 0; k--) if (v->voies_ok[k]) { v->voie_desiree = nV - 1; return; }   // one lane per time step (no reserved lanes in scope)
      for (int k = nV + 2; k < n; k++) if (v->voies_ok[k]) { v->voie_desiree = nV + 1; return; }
    }
  }
  Veh* get_first_vehicule_lane(int lane) {                                                   // reseau.cpp:3255-3295
    Veh* r = nullptr; double p = 0;
    for (Veh* c : lane_set[lane]) if (p <= c->pos[1]) { r = c; p = c->pos[1]; }
    return r;
  }
  Veh* get_first_vehicule(std::vector<Veh*>& l) {                                            // reseau.cpp:14084-14128
    if (l.size() == 1) return l[0];
    if (l.empty()) return nullptr;
    n_ties++;
    for (size_t i = 0; i < l.size(); i++) { Veh* ld = get_leader_of(l[i]); if (!ld) return l[i];
      bool found = false; for (size_t j = 0; j < l.size() && !found; j++) if (i != j && ld->id == l[j]->id) found = true;
      if (!found) return l[i]; }
    return l.back();
  }
  Veh* get_near_amont_vehicule(int lane, double pos, double dstmax, Veh* excl) {             // reseau.cpp:3374-3520
    std::vector<Veh*> l; if (lane < 0) return nullptr;
    const int lane0 = lane;                                                                  // pTuyau = parent of the lane the search started on (:3411)
    if (dstmax < 0) dstmax = 9999;
    double ptmp = -99999;
    for (Veh* c : lane_set[lane]) if (pos > c->pos[1] && c->pos[1] >= ptmp) { if (!excl || excl->id != c->id) {
      if (c->pos[1] == ptmp) l.push_back(c); else { l.resize(1); l[0] = c; ptmp = c->pos[1]; } } }
    if (dstmax > 0 && !l.empty() && (pos - ptmp) < dstmax) return get_first_vehicule(l);
    dstmax -= pos; ptmp = -99999;
    int pv = lanes[lane].prev_lane;                                                          // VoieMicro::GetPrevVoie
    if (l.empty() && pv >= 0 && dstmax > 0) {
      pos = lanes[pv].length + 0.001;
      for (Veh* c : lane_set[pv]) if (pos >= c->pos[1] && c->pos[1] >= ptmp) { if (!excl || excl->id != c->id) {
        if (c->pos[1] == ptmp) l.push_back(c); else { l.resize(1); l[0] = c; ptmp = c->pos[1]; } } }
      lane = pv;
    }
    if (!l.empty()) { if ((pos - ptmp) < dstmax) return get_first_vehicule(l); else return nullptr; }
    // :3479-3517 the upstream connection has a single upstream link with a single lane (lcam = its m_LstTuyAm; empty without a connection)
    int link0 = lanes[lane0].link;
    dstmax -= pos;
    if (dstmax > 0 && lcam[link0].size() == 1 && links[lcam[link0][0]].n_lanes == 1) {
      int up = links[lcam[link0][0]].first_lane; pos = lanes[up].length + 0.001; ptmp = -99999;
      for (Veh* c : lane_set[up]) if (pos >= c->pos[1] && c->pos[1] >= ptmp) { if (!excl || excl->id != c->id) {
        if (c->pos[1] == ptmp) l.push_back(c); else { l.resize(1); l[0] = c; ptmp = c->pos[1]; } } }
    }
    return get_first_vehicule(l);
  }
  double distance_entre_vehicules(Veh* a, Veh* b) {                                          // reseau.cpp:4446-4560 (start of step)
    if (!a || !b) return 9999;
    if (a->link[1] == b->link[1]) return fabs(b->pos[1] - a->pos[1]);
    if (a->link[1] < 0) { if (a->lane[1] == b->lane[1]) return b->pos[1];
      if (links[lanes[a->lane[1]].link].down_node == links[b->link[1]].up_node) return lanes[a->lane[1]].length + b->pos[1]; return 9999; }
    if (b->link[1] < 0) { if (b->lane[1] == a->lane[1]) return a->pos[1];
      if (b->lane[1] >= 0 && links[lanes[b->lane[1]].link].down_node == links[a->link[1]].up_node) return lanes[b->lane[1]].length + a->pos[1]; return 9999; }
    if (links[a->link[1]].down_node == links[b->link[1]].up_node) return lanes[a->lane[1]].length - a->pos[1] + b->pos[1];
    if (links[b->link[1]].down_node == links[a->link[1]].up_node) return lanes[b->lane[1]].length - b->pos[1] + a->pos[1];
    return 9999;
  }
  // CDiagrammeFondamental::CalculVitEqDiagrOrig(concentration) DiagFonda.cpp:116-172, :221-276 (bFluide = true)
  static double vit_eq(double w, double kx, double vx, double conc) {
    double kc = w * kx / (w - vx); double debit;
    if (conc <= kc) { debit = vx * conc; if (fabs(debit) < DBL_EPSILON) debit = 0; }
    else { debit = w * (conc - kx); if (fabs(debit) < DBL_EPSILON) debit = 0; }
    if (fabs(debit) < DBL_EPSILON && fabs(conc) < DBL_EPSILON) return 0;
    double vit = conc < DBL_EPSILON ? vx : debit / conc;
    if (fabs(vit) < DBL_EPSILON) vit = 0;
    return vit;
  }
  // AbstractCarFollowing::TestLaneChange AbstractCarFollowing.cpp:61-329 (no pull-back, no ghost, synchronous mandatory mode)
  bool test_lane_change(Veh* v, Veh* fol, Veh* lead, Veh* lead_orig, int tgt_lane, bool force) {
    const Cls& c = cls[v->cls]; double kc = c.k_crit();
    double si;
    if (lead_orig) si = lead_orig->pos[1] - v->pos[1];
    else if (force) si = std::max<double>(1 / v->kx, lanes[v->lane[1]].length - v->pos[1]);
    else si = 1 / kc;
    double sf = (fol && lead) ? distance_entre_vehicules(lead, fol) : 1 / kc;
    double vit = std::min<double>(v->vit_reg_tuy_act, v->vx);
    double dem_orig = std::min<double>(vit / si, c.debit_max());                             // :130-146
    double dem_dest;
    if (fol) dem_dest = std::min<double>(fol->vx / sf, cls[fol->cls].debit_max());          // :149-153
    else dem_dest = std::min<double>(vit / sf, c.debit_max());
    double off_dest = lead ? std::min<double>((lead->kx - (1 / sf)) * (-1) * lead->w, cls[lead->cls].debit_max()) : c.debit_max();   // :158-161
    double pi;
    if (force) pi = 1 / dt;                                                                  // :168-173
    else {
      double vtgt;
      if (sf > 0) { const Veh* d = lead ? lead : v; vtgt = vit_eq(d->w, d->kx, d->vx, 1 / sf); if (lead) vtgt = lead->vit[1]; } else vtgt = 0;   // :177-194
      vtgt = std::min<double>(vtgt, links[lanes[tgt_lane].link].vit_reg);
      if (!fol && lead) { vtgt = std::max<double>(vtgt, lead->vit[1]); vtgt = std::min<double>(vtgt, lead->vit_reg_tuy_act); }
      vtgt = std::min<double>(vtgt, v->vit_max);
      double vorg = DBL_MAX;
      if (si > 0) { const Veh* d = lead_orig ? lead_orig : v; vorg = vit_eq(d->w, d->kx, d->vx, 1 / si);
        if (c.law == 0 && v->ctx && v->ctx->deltaN < 1 && lead_orig) vorg = lead_orig->vit[1]; }   // :218-222
      else vorg = lead_orig ? lead_orig->vit[1] : v->vit[1];
      vorg = std::min<double>(vorg, links[v->link[1]].vit_reg); vorg = std::min<double>(vorg, v->vit_max);
      pi = std::max<double>(0., vtgt - vorg) / (links[v->link[1]].vit_reg * c.tau_lc);        // :246
      if (pi < 0.000001) return false;
    }
    double phi;
    if (force && links[v->link[1]].length - v->pos[1] < dst_chgtvoie_force) phi = phi_chgtvoie_force;   // :254-257
    else {
      if (!fol || !lead) dem_dest = 0;                                                       // :260-261
      if (dem_dest > 0) phi = std::min<double>(1., off_dest / dem_dest) * pi * dem_orig / vit * dt * si;
      else phi = pi * dem_orig / vit * dt * si;
    }
    double r = rnd.myRand() / MAXIMUM_RANDOM_NUMBER;                                         // :324
    return r < phi;
  }
  void changement_voie(Veh* v, int tgt_lane, Veh* fol) {                                    // vehicule.cpp:2657-2711
    v->b_chgt_voie = true;
    int old = v->lane[1];
    set_voie_ant(v, tgt_lane);
    v->pos[1] = v->pos[1] * lanes[tgt_lane].length / lanes[old].length;
    if (fol && fol->link[1] == v->link[1]) v->pos[1] = std::max<double>(v->pos[1], fol->pos[1]);
    v->voie_desiree = -1;
    v->next_voie = calcul_next_voie(v, v->lane[1]);                                          // :2700 (consumes the lane's memoised draw; m_pNextVoie is kept)
  }
  bool calcul_changement_voie(Veh* v, int tgt_num, bool force) {                            // vehicule.cpp:2248-2457
    if (v->pos[1] <= UNDEF_POSITION) return false;                                           // :2291
    if (v->b_chgt_voie) return false;                                                        // :2297
    int lane = v->lane[1]; bool desired = v->voie_desiree != -1;
    if (get_first_vehicule_lane(lane) == v && !chgt_voie_obligatoire(lane) && !desired) return false;   // :2302-2307
    int tgt = links[v->link[1]].first_lane + tgt_num;
    double ratio_pos = lanes[tgt].length / lanes[lane].length * v->pos[1];
    bool move_back = false;
    Veh* same = get_near_aval_vehicule(tgt, ratio_pos - 0.01, v);                            // :2378
    if (same && fabs(lanes[tgt].length / lanes[lane].length * v->pos[1] - same->pos[1]) < 0.01) {
      if (fabs(v->pos[1] - lanes[lane].length) < 0.01) {
        if (force && same->voie_desiree == lanes[lane].num && v->voie_desiree == lanes[same->lane[1]].num && v->vit[1] < 0.01 && same->vit[1] < 0.01) {
          changement_voie(v, tgt, nullptr); changement_voie(same, lane, nullptr); return true;          // switch-switch :2383-2393
        }
        v->pos[1] = v->pos[1] - 0.5; move_back = true;
      } else return false;
    }
    double q = lanes[tgt].length / lanes[lane].length * v->pos[1] + 0.001;
    Veh* fol = get_near_amont_vehicule(tgt, q, -1.0, v);                                     // :2403
    Veh* lead = get_near_aval_vehicule(tgt, q, v);                                           // :2404
    if (!lead && (chgt_voie_obligatoire(lane) || desired)) {                                 // :2417-2434
      bool cont = true;                                                                      // non-priority branch of a merge: no leader looked for downstream
      if (links[v->link[1]].type_aval == 'C' && link_cvg_aval[v->link[1]] >= 0 && voie_amont_non_prioritaire(cvgs[link_cvg_aval[v->link[1]]], lane)) cont = false;
      if (cont) { int nx = calcul_next_voie(v, tgt); if (nx >= 0) lead = get_near_aval_vehicule(nx, 0, nullptr); }
    }
    Veh* lead_orig = get_prev_vehicule(v);                                                   // :2439
    if (!lead_orig && !chgt_voie_obligatoire(lane) && !desired && !force) return false;   // :2440-2442 (the reference returns here without undoing the 0.5 m move-back)
    if (test_lane_change(v, fol, lead, lead_orig, tgt, force)) { changement_voie(v, tgt, fol); return true; }
    if (move_back) v->pos[1] = v->pos[1] + 0.5;
    return false;
  }
  void changement_de_voie() {                                                                // reseau.cpp:4219-4330
    for (Veh* v : lst) v->b_chgt_voie = false;
    for (Veh* v : lst) if (v->lane[1] >= 0 && v->voie_desiree != -1) {                       // mandatory pass
      bool ok = calcul_changement_voie(v, v->voie_desiree, true);
      v->b_chgt_voie = true;
      if (ok) v->voie_desiree = -1;
    }
    for (Veh* v : lst) if (v->lane[1] >= 0 && !v->b_chgt_voie && v->link[1] >= 0) {          // discretionary pass
      int nV = lanes[v->lane[1]].num, n = links[v->link[1]].n_lanes;
      int nVF = nV + 1, nVS = nV - 1;
      if (ordre_chgtvoie == 2) { nVF = nV - 1; nVS = nV + 1; }
      if (nVF >= 0 && nVF < n && v->voies_ok[nVF] && calcul_changement_voie(v, nVF, false)) continue;
      if (nVS >= 0 && nVS < n && v->voies_ok[nVS] && calcul_changement_voie(v, nVS, false)) continue;
      // pull-back (:4320-4326) only from reserved lanes: none in scope
    }
  }

  // ------------------------------------------------------------------------------------------
  // context: CarFollowingContext.cpp:38-170, NewellContext.cpp:31-169, AbstractCarFollowing.cpp:671-685
  void build_context(Veh* v, double range) {
    v->prev_ctx = std::move(v->ctx);
    v->ctx = std::make_unique<Ctx>(); Ctx& c = *v->ctx; Ctx* prev = v->prev_ctx.get();
    c.start_pos = v->pos[1];
    // BuildLstLanes :108-125
    int lane = v->lane[1]; c.lanes.push_back(lane);
    double dist = len_ex(lane) - (v->pos[1] + v->creation_pos);
    while (dist < range && (lane = calcul_next_voie(v, lane)) >= 0) { c.lanes.push_back(lane); dist += len_ex(lane); }
    // SearchLeaders :128-170 (the `break` leaves the lane loop at the first vehicle found, accepted or not)
    double d = 0;
    for (size_t i = 0; i < c.lanes.size(); i++) {
      int ln = c.lanes[i];
      if (i == 0) {
        Veh* ld = get_prev_vehicule(v);
        if (ld) { double g = ld->pos[1] - (v->pos[1] + v->creation_pos); if (g <= range) { c.leader = ld; c.leader_id = ld->id; c.leader_dist = g; } break; }
        d += len_ex(ln) - (v->pos[1] + v->creation_pos);
      } else {
        Veh* ld = get_near_aval_vehicule(ln, 0, nullptr);
        if (ld) { double g = ld->pos[1] + d; if (g <= range) { c.leader = ld; c.leader_id = ld->id; c.leader_dist = g; } break; }
        d += len_ex(ln);
      }
    }
    if (cls[v->cls].law == 0) {                                                              // NewellContext::Build :31-44
      if (prev) c.contraction = prev->contraction;
      // ComputeDeltaN :121-169
      double dn;
      if (!c.leader) dn = 1;
      else if (c.contraction) dn = prev->deltaN;
      else if (!prev) dn = 1;
      else if (prev->leader_id < 0 || prev->leader_id != c.leader_id || prev->free_flow) {
        double kv = -c.leader->kx * c.leader->w / (c.leader->vit[1] - c.leader->w);
        dn = std::min<double>(1, kv * c.leader_dist);
      } else dn = prev->deltaN;
      c.deltaN = dn;
    }
  }

  // ------------------------------------------------------------------------------------------
  // Newell: NewellCarFollowing.cpp:53-136, :204-398
  double newell_free(Veh* v, double dtv, const Ctx& c) {                                    // :70-111
    double start = c.start_pos, trav = 0, rem = dtv;
    for (size_t i = 0; i < c.lanes.size(); i++) {
      int ln = c.lanes[i]; double sp = i == 0 ? start : 0; double lex = len_ex(ln);
      double vmax_lane = vit_reg(lanes[ln].link);
      double vdes = std::min<double>(v->vit_max, vmax_lane);
      if (accbornee) vdes = std::min<double>(vdes, v->vit[1] + cls[v->cls].acc_max(v->vit[1]) * dtv);
      double tt = (lanes[ln].length - sp) / vdes;
      if (tt >= rem || i == c.lanes.size() - 1) { trav += rem * vdes; break; }
      trav += lex - sp; rem -= tt;
    }
    return trav;
  }
  double newell_relax(double cur, Veh* v, Veh* L, double vL, double dtv, const Ctx& c) {    // :113-136
    double kv = -L->kx * L->w / (L->vit[1] - L->w);
    double kvfin = -L->kx * L->w / (vL - L->w);
    double r = cur / (-v->kx * L->w);
    if (c.contraction) return cur;
    double res;
    if ((dt - r) >= -0.0001) res = std::min<double>(1, (cur / kv + std::min<double>(r / dtv * (vL - L->vit[1]) + relax, vL) * dt) * kvfin);
    else res = std::min<double>(1, (cur / kv + std::min<double>((vL - L->vit[1]) + relax, vL) * dt) * kvfin);
    return std::max<double>(cur, res);
  }
  double newell_cong(Veh* v, double dtv, Ctx& c) {                                          // :204-398
    Veh* L = c.leader; double dist = c.leader_dist; double vL;
    if (L->deja_calcule && same_part(v, L)) vL = L->vit[0];
    else {                                                                                   // :217-230 (at most one leader in the context)
      if (accbornee) vL = std::min<double>(std::min<double>(L->vit_reg_tuy_act, L->vit_max), L->vit[1] + cls[v->cls].acc_max(L->vit[1]) * dt);
      else vL = std::min<double>(L->vit_reg_tuy_act, L->vit_max);
    }
    // :233-263 red-light correction: the inner `ControleurDeFeux *pCDF` declarations shadow the outer
    // NULL pointer, so the branch never fires — nothing to restate.
    // :266-298 the leader heads a non-priority merge branch: worst case, it does not insert
    if (!c.post && L->link[1] >= 0 && links[L->link[1]].type_aval == 'C' && link_cvg_aval[L->link[1]] >= 0) {
      const Cvg& cv = cvgs[link_cvg_aval[L->link[1]]];
      if (cv.am_links.size() >= 1 && get_prev_vehicule(L) == nullptr && voie_amont_non_prioritaire(cv, L->lane[1])) {
        double vit = -1;
        if (!L->deja_calcule) vit = calcul_vitesse_leader(cv, L, inst_cur, dtv); else vit = L->vit[0];
        const double borne = std::min<double>(L->vit[1] + cls[L->cls].acc_max(v->vit[1]) * dtv, L->vit_reg_tuy_act);   // sic: the leader's table at the follower's speed
        if (vit == -1) vit = borne; else vit = std::min<double>(vit, borne);
        if (L->pos[1] + vit * dtv > lanes[L->lane[1]].length) vL = (lanes[L->lane[1]].length - L->pos[1]) / dtv;
      }
    }
    double dn0 = c.deltaN;
    double dn1 = newell_relax(dn0, v, L, vL, dtv, c);
    c.deltaN = dn1;
    double r = dn0 / (-L->kx * L->w);
    double kvfin = -L->kx * L->w / (vL - L->w);
    double pc;
    if ((dt - r) >= -0.0001) {
      pc = dist + vL * dt - dn1 / kvfin;
      if (pc <= 0.0001 && dn0 < 1) { c.deltaN = dn0; pc = dist + vL * dt - dn0 / kvfin; }
    } else {
      double alpha = -L->w * L->kx * dt / dn0;
      if (dtv < dt) pc = dist + vL * dt - dn1 / kvfin;
      else { pc = alpha * dist + vL * dt - alpha * dn1 / kvfin;
        if (pc <= 0.0001 && dn0 < 1) { c.deltaN = dn0; pc = alpha * dist + vL * dt - alpha * dn0 / kvfin; } }
    }
    return pc;
  }
  // Gipps: GippsCarFollowing.cpp:50-122 ; IDM: IDMCarFollowing.cpp:61-211 ; distance AbstractCarFollowing.cpp:692-716
  // std::pow is the host libm's (glibc: within 0.52 ulp, i.e. not the correctly rounded value for ~0.08 % of the arguments; another
  // platform's libm differs again).  ORACLE_POW=cr swaps in the correctly rounded value of the three powers the laws use, the
  // same double-double evaluation as the device (csrc/micro.cuh), so that the tests can separate the engine's logic (bit-exact
  // under identical arithmetic) from the libm sensitivity of Gipps / IDM (1 ulp, amplified by the car-following feedback).
  bool pow_cr = [] { const char* e = getenv("ORACLE_POW"); return e && !strcmp(e, "cr"); }();   // read when the simulation is created
  double pow_law(double x, double y) const {
    if (!pow_cr) return pow(x, y);
    if (y == 0.5) return sqrt(x);
    if (y == 4) { double p1 = x * x, e1 = std::fma(x, x, -p1); double p2 = p1 * p1, e2 = std::fma(p1, p1, -p2); e2 = e2 + 2.0 * p1 * e1; return p2 + e2; }
    if (y == 1.5 && x > 0 && x <= DBL_MAX) { double r0 = sqrt(x); double r = std::fma(-r0, r0, x); double c = r / (2.0 * r0); double p = x * r0, e = std::fma(x, r0, -p); e = e + x * c; return p + e; }
    return pow(x, y);
  }
  double compute_law(Veh* v, double dtv, Ctx& c) {
    const Cls& k = cls[v->cls];
    if (k.law == 0) {                                                                        // Newell InternalCompute :53-68
      bool ff = true; double d = newell_free(v, dtv, c);
      if (c.leader) { double dc = newell_cong(v, dtv, c); if (dc < d) { ff = false; d = dc; } }
      c.free_flow = ff; return d;
    }
    int fl = c.lanes.front(); double vmaxl = std::min<double>(v->vit_max, vit_reg(lanes[fl].link));
    if (k.law == 1) {
      const double A = 2.5, B = 0.025, G = 0.5, TH = 1;
      bool ff = true;
      double vf = v->vit[1] + A * k.acc_max(v->vit[1]) * dtv * (1 - v->vit[1] / vmaxl) * pow_law(B + v->vit[1] / vmaxl, G);
      if (c.leader) { Veh* L = c.leader;
        double dmax = -k.decel; if (dmax == 0) dmax = -DBL_MAX;
        double dlmax = -cls[L->cls].decel; if (dlmax == 0) dlmax = -DBL_MAX;
        double fac = dmax * (dtv / 2.0 + TH);
        double sq = fac * fac - dmax * (2.0 * (c.leader_dist - (veh_length(L) + k.esp)) - v->vit[1] * dtv - L->vit[1] * L->vit[1] / dlmax);
        double vc = sq >= 0 ? fac + sqrt(sq) : 0;
        if (vc < vf) { ff = false; vf = vc; } }
      c.free_flow = ff; return vf * dtv;
    }
    // IDM (non-improved: CarFollowingFactory.cpp:118)
    double sp = v->vit[1], relpos, relsp;
    if (c.leader) { relpos = c.leader_dist - veh_length(c.leader); relsp = v->vit[1] - c.leader->vit[1]; }
    else { relpos = DBL_MAX; relsp = v->vit[1]; }
    double amax = k.acc_max(sp); double dec = k.decel; if (dec == 0) dec = DBL_MAX;
    double z;
    if (relpos <= 0) z = 2;
    else { double dsp = k.esp + std::max<double>(1 * sp + (sp * relsp) / (2.0 * sqrt(amax * dec)), 0); z = dsp / relpos; }
    c.free_flow = z < 1;
    double a = amax * (1.0 - pow_law(sp / vmaxl, 4) - pow_law(z, 1.5));
    return v->vit[1] * dtv + 0.5 * a * dtv * dtv;
  }

  // ------------------------------------------------------------------------------------------
  // Vehicule::CalculDeceleration vehicule.cpp:1017-1074
  double calcul_deceleration(Veh* v, double pas, double limite) {
    double g = limite;
    if (v->ctx->free_flow) {
      v->b_deceleration = true;
      if (v->deceleration == -1) v->deceleration = cls[v->cls].decel;                       // vehicule.cpp:142-154 (sigma = 0)
      double dd = 0;
      if (v->deceleration != 0) { double N = floor(v->vit[1] / (v->deceleration * dt)); dd = N * v->vit[1] * dt - ((N + 1) * N / 2) * v->deceleration * dt * dt; }
      else return g;
      if (v->vit[1] == 0) { double tm = sqrt(2.0 * limite / v->deceleration); double s = v->deceleration * tm; g = std::min<double>(g, s * pas); }
      else { double ld = limite - dd; double T = ld / v->vit[1];
        if (T <= pas) { if (T < 0) T = 0; g = std::max<double>(std::min<double>(g, v->vit[1] * pas - v->deceleration * (pas - T)), 0); } }
    }
    return g;
  }
  // Vehicule::CalculTraficFeux vehicule.cpp:1799-1945
  bool calcul_trafic_feux(Veh* v, double inst, double dtv, double start_of_lane, double start_on_lane, double end_on_lane,
                          double goal, int lane, size_t ilane, const std::vector<int>& lst, double& mini, bool& arret) {
    bool infl = false; int aval = -1;
    for (auto& lf : lanes[lane].sl) {
      if (lf.pos >= start_on_lane) {
        int am = links[lanes[lane].link].brick >= 0 ? lanes[lane].prev_lane : lane;          // :1840-1847 (GetPrevVoie of an internal lane)
        if (am >= 0 && lf.up_link == lanes[am].link && lf.up_num == lanes[am].num) {
          if (aval < 0) {
            for (size_t n = ilane + 1; n < lst.size() && aval < 0; n++) if (links[lanes[lst[n]].link].brick < 0) aval = lst[n];
            if (aval < 0) { aval = lst.back(); do { aval = calcul_next_voie(v, aval); } while (aval >= 0 && links[lanes[aval].link].brick >= 0); }
          }
          if (aval >= 0 && lf.down_link == lanes[aval].link && lf.down_num == lanes[aval].num) {
            double p = 0; int col = statut_couleur(lf.ctrl, inst, lf.up_link, lf.down_link, &p);
            double dloc = (start_of_lane - start_on_lane) + lf.pos;
            if (col == 0 && fabs(p) < 0.00001) { mini = dloc; arret = true; return true; }
            double ts = v->vit[1] > 0 ? dloc / v->vit[1] : dloc / (end_on_lane / dtv);
            if (col != 1 && p > 0) { if (ts > p * dtv) { mini = dloc; arret = true; return true; } }
            if (col == 1 && p < 1) {
              if (ts > dtv * (1 - p)) {} else { infl = true; if (v->vit[1] > 0) mini = goal - (1 - p) * v->vit[1] * dtv; else mini = goal * p; }
            }
            break;
          }
        }
      }
    }
    return infl;
  }
  // Sortie::GetNextEnterInstant sortie.cpp:60-118
  double next_enter_instant(const Exit& x, int nvoie, double prev, double inst, double pas) {
    double cap = x.capacity < 0 ? NONVAL_DOUBLE : x.capacity;
    if (cap >= NONVAL_DOUBLE) return inst;
    double cum = 0, cump = 0, ti = prev;
    while (cum < 1 && ti <= duree) { cump = cum; cum += cap * pas / nvoie; ti += pas; }
    if (cum < 1) return -1;
    return (ti - pas) + (1 - cump) / (cum - cump) * pas;
  }
  // Vehicule::CheckNetworkInfluence vehicule.cpp:1591-1708
  bool check_network_influence(Veh* v, double inst, double dtv, int lane, double llen, size_t ilane, const std::vector<int>& lst,
                               double trav, double goal, double start, double end, double& mini) {
    bool infl = false; mini = DBL_MAX; const Link& T = links[lanes[lane].link];
    if (end > llen) {                                                                        // 1 - lane end
      bool stop = chgt_voie_obligatoire(lane) || ((ilane == lst.size() - 1) && T.type_aval != 'S' && T.type_aval != 'P');
      if (stop) { v->regime_fluide = false; infl = true; mini = std::min<double>(mini, trav + llen - start); }
    }
    // 2 - next trip node deceleration: the only trip node in scope is the final exit, reached by crossing
    // the lane end (GetNextTripNodeDistance returns DBL_MAX for a punctual exit) -> no effect.
    if (T.type_aval == 'S' || T.type_aval == 'P') {                                          // exit capacity :1636-1666
      double pcrit = llen - max_influence(v);
      if (end > pcrit) {
        double ts = lanes[lane].next_inst_sortie;
        if (ts < 0) { infl = true; mini = std::min<double>(mini, trav + pcrit - start); }
        else if (ts > inst - dtv) { double vv = (trav + llen - start) / (ts - (inst - dtv)); infl = true; mini = std::min<double>(mini, vv * dtv); }
      }
    }
    // 3 - downstream punctual connection: Convergent::ComputeTraffic convergent.cpp:1368-1403 (vehicule.cpp:1668-1677)
    { int cv = link_cvg_aval[lanes[lane].link];
      if (cv >= 0 && lane == v->lane[1] && voie_amont_non_prioritaire(cvgs[cv], lane) && get_prev_vehicule(v) == nullptr) {
        double mc = calcul_vitesse_leader(cvgs[cv], v, inst, dtv) * dtv;
        if (mc < goal) { infl = true; mini = std::min<double>(mini, mc); } } }
    // 4 - signals :1680-1705
    double ml = goal + 1; bool arret = false;
    if (calcul_trafic_feux(v, inst, dtv, trav, start, end, goal, lane, ilane, lst, ml, arret)) {
      if (arret) { double md = calcul_deceleration(v, dtv, ml); ml = std::min<double>(ml, md); }
      if (ml < goal) { infl = true; if (ml < mini) { mini = ml; if (arret) { v->regime_fluide = false; v->arret_au_feu = true; } } }
    }
    return infl;
  }

  bool is_leg_done(Veh* v, int lane, double llen, double end) {                             // TripLeg.cpp:91-117, Position.cpp:394-404
    const Link& T = links[lanes[lane].link];
    if (!(end > llen && T.type_aval == 'S')) return false;
    return routes[v->route].back() == lanes[lane].link;
  }
  void update_used_lanes(Veh* v) {                                                          // vehicule.cpp:1751-1775
    if (!v->used_lanes.empty()) {
      for (size_t i = 0; i < v->used_lanes.size(); i++) if (v->used_lanes[i] == v->lane[0]) { v->used_lanes.resize(i); break; }
      if (v->link[0] >= 0 && v->link[0] == v->link[1]) v->used_lanes.clear();
    }
  }
  void set_voie_ant(Veh* v, int lane) {                                                     // vehicule.cpp:928-955
    if (v->lane[1] >= 0) lane_set[v->lane[1]].erase(v);
    v->lane[1] = lane;
    if (lane >= 0) lane_set[lane].insert(v);
  }
  static void set_acc(Veh* v, double a) { if (fabs(a) < 0.000001) a = 0; v->acc[0] = a; }  // vehicule.h:501

  // ------------------------------------------------------------------------------------------
  // Vehicule::CalculTraficEx vehicule.cpp:1192-1588
  void calcul_trafic_ex(Veh* v, double inst, std::vector<int>& ids) {
    if (v->deja_calcule) return;
    for (int id : ids) if (id == v->id) { if (cls[v->cls].law == 0) n_loops++; return; }   // :1203-1216
    ids.push_back(v->id);
    if ((long)ids.size() > max_depth) max_depth = (long)ids.size();
    int prevlane = v->lane[1];
    if (prevlane < 0) { prevlane = v->lane[0]; set_voie_ant(v, prevlane); }                 // :1220-1232
    bool created;
    if (v->pos[1] <= UNDEF_POSITION) {                                                       // :1235-1268 (generation position never > 0 at an Entree)
      created = true;
      if (v->pos[0] <= UNDEF_POSITION) v->pos[1] = 0; else v->pos[1] = v->pos[0];
    } else { created = false; v->creation_pos = 0; }
    double dtv = dt - std::max<double>(0, v->inst_creation);                                // :1278
    double range = max_influence(v) + dtv * v->vit_max;                                     // :1299, :1710-1715
    build_context(v, range);                                                                 // :1302
    Ctx& c = *v->ctx; inst_cur = inst;
    if (c.leader && same_part(v, c.leader)) calcul_trafic_ex(c.leader, inst, ids);          // :1304-1309 leader first
    double goal = compute_law(v, dtv, c);                                                    // :1312
    v->regime_fluide = c.free_flow; v->regime_fluide_leader = v->regime_fluide;             // :1313-1315
    v->arret_au_feu = false;
    double start = v->pos[1] + v->creation_pos;
    double mgoal = goal + 1, trav = 0;
    const std::vector<int> lst = c.lanes;
    for (size_t i = 0; i < lst.size(); i++) {                                               // :1324-1341
      int ln = lst[i]; double sp = i == 0 ? start : 0; double llen = len_ex(ln); double ep = goal - trav + sp; double ml;
      if (check_network_influence(v, inst, dtv, ln, llen, i, lst, trav, goal, sp, ep, ml)) if (ml < mgoal) mgoal = ml;
      trav += std::min<double>(ep, llen) - sp;
    }
    goal = std::min<double>(goal, mgoal);                                                    // :1343
    bool cancel = false;
    if (created && goal + v->creation_pos <= 0) { cancel = true; v->creation_pos = v->creation_pos + goal; }   // :1346-1351
    goal = std::max<double>(0, goal);
    if (created && !v->regime_fluide && c.leader) v->vit[0] = std::min<double>(v->vit_max, c.leader->vit[1]);   // :1358-1365
    else v->vit[0] = dtv != 0 ? goal / dtv : v->vit_max;
    if (created) set_acc(v, 0); else set_acc(v, dtv != 0 ? (v->vit[0] - v->vit[1]) / dtv : 0);   // :1368-1375
    v->used_lanes.clear(); trav = 0; double trav_dest = 0; int dest_lane = -1;
    for (size_t i = 0; i < lst.size(); i++) {                                               // :1384-1506
      int ln = lst[i]; double sp = i == 0 ? start : 0; double llen = len_ex(ln); double ep = goal - trav + sp;
      trav += std::min<double>(ep, llen) - sp;
      char ta = links[lanes[ln].link].type_aval; bool is_exit = ta == 'P' || ta == 'S'; bool last = i == lst.size() - 1;
      if ((last && !is_exit && ep > llen) || (v->arret_au_feu && ep > llen && ep < llen + 0.00001)) ep = llen;   // :1397-1403
      bool reached = is_leg_done(v, ln, llen, ep);
      if (reached) { dest_lane = ln; trav_dest = trav; }
      if (reached) { v->link[0] = lanes[ln].link; v->lane[0] = ln; v->pos[0] = ep; break; }  // :1417-1423
      else if (ep <= llen) { v->link[0] = lanes[ln].link; v->lane[0] = ln; v->pos[0] = ep; break; }   // :1425-1431
      { int cv = link_cvg_aval[lanes[ln].link]; if (cv >= 0 && ta != 'E' && ta != 'S') cvgs[cv].ins_veh.push_back(v); }   // :1485-1496 AddInsVeh
      if (i > 0) v->used_lanes.push_back(ln);                                                // :1499-1502
      v->voie_desiree = -1;                                                                  // :1505
    }
    if (dest_lane >= 0) {                                                                    // :1512-1527
      double ti = goal > 0.0001 ? inst - dtv + dtv * (trav_dest / goal) : inst - dtv;
      v->end_leg = true; v->dest_enter_lane = dest_lane; v->inst_dest_reached = ti;
    }
    if (!v->end_leg && v->link[0] >= 0) update_used_lanes(v);                               // :1560-1564 (traversees off)
    if (cancel) { v->pos[1] = UNDEF_POSITION; v->pos[0] = UNDEF_POSITION; }                 // :1567-1571
    v->deja_calcule = true; v->inst_creation = 0;                                           // :1573-1574
  }

  // ------------------------------------------------------------------------------------------
  // Vehicule::DecalVarTrafic vehicule.cpp:958-1015 (history depth > 2 is never read by the micro path)
  void decal(Veh* v) {
    v->deja_calcule = false; v->attente_insertion = false; v->b_deceleration = false; v->end_leg = false;
    if (v->lane[0] != v->lane[1]) v->prev_lane = v->lane[1];
    v->pos[1] = v->pos[0]; if (v->pos[1] != UNDEF_POSITION && fabs(v->pos[1]) < 0.00001) v->pos[1] = 0;
    v->lane[1] = v->lane[0]; v->link[1] = v->link[0]; v->vit[1] = v->vit[0]; v->acc[1] = v->acc[0];
    v->lane[0] = -1; v->link[0] = -1; v->pos[0] = 0; v->vit[0] = 0; v->acc[0] = 0;
  }
  // Reseau::InitVehicules reseau.cpp:3901-3976
  void init_vehicules() {
    for (auto& s : lane_set) s.clear();
    for (Veh* v : lst) {
      decal(v);
      if (v->lane[1] >= 0) lane_set[v->lane[1]].insert(v);
      if (v->link[1] >= 0) {
        if (v->next_link < 0) v->next_link = calcul_next_tuyau(v, v->link[1]);
        calcul_voies_possibles(v);                                                           // :3951
        calcul_voie_desiree(v);                                                              // :3953
      }
    }
  }

  // ------------------------------------------------------------------------------------------
  // demand: usage/SymuViaTripNode.cpp:847-955 (UpdateCritCreationVeh), :212-325 (CreationVehicules),
  // :331-840 (GenerateVehicle), :1642-1851 (InsertionVehEnAttente)
  double demande_value(const Entry& e, double inst) const {
    double acc = 0; for (auto& d : e.demand) { acc += d.first; if (inst <= acc) return d.second; }
    return e.demand.empty() ? 0 : e.demand.back().second;
  }
  void update_crit_creation(Entry& e, double inst, bool due_to_creation) {                  // :847-955
    if (!e.exponential) { e.crit += demande_value(e, inst) * dt; return; }                  // :905-911
    if (fabs(demande_value(e, inst) - demande_value(e, inst - dt)) > DBL_EPSILON || fabs(inst - dt) < DBL_EPSILON || due_to_creation) {   // :915-917
      double niveau = demande_value(e, inst); if (niveau > max_debit_max) niveau = max_debit_max;
      double r = 0; while (r <= 0) r = (double)rnd.myRand() / MAXIMUM_RANDOM_NUMBER;
      e.crit = niveau > 0 ? inst - log(r) * (1 - niveau / max_debit_max) / niveau + 1 / max_debit_max : DBL_MAX;   // :935-940
    }
  }
  Veh* generate_vehicle(Entry& e, double inst, double inst_creat) {                         // :331-840
    const Link& T = links[e.link];
    // lane draw: CalculNumVoie :1928-1999
    int nv = 0; { double r = rnd.myRand() / MAXIMUM_RANDOM_NUMBER; double sum = 0;
      for (; nv < T.n_lanes; nv++) { double cf = e.lane_coefs[nv]; if (cf <= 0) continue; sum += cf; if (sum >= r) break; }
      if (nv >= T.n_lanes) nv = T.n_lanes - 1; }
    // type draw: RepartitionTypeVehicule.cpp:78-125
    int tv = -1; { double r = rnd.myRand() / MAXIMUM_RANDOM_NUMBER; double sum = 0; int lastnn = -1;
      for (size_t i = 0; i < cls.size(); i++) { double cf = e.type_coefs[i]; if (cf <= 0) continue; lastnn = (int)i; sum += cf; if (r <= sum) { tv = (int)i; break; } }
      if (tv < 0) tv = lastnn; }
    Veh* v = new_vehicle(tv, last_id++);                                                     // IncLastIdVeh, ctor, InitializeCarFollowing
    // destination draw: GetDestination :1255-1322
    int d = -1; { double r = (double)rnd.myRand() / MAXIMUM_RANDOM_NUMBER; double sum = 0;
      for (size_t i = 0; i < e.dests.size(); i++) { if (e.dests[i].coeff <= 0) continue; sum += e.dests[i].coeff; if (i == e.dests.size() - 1) sum = 1;
        if (r <= sum) { d = (int)i; break; } } }
    if (d < 0) { return nullptr; }
    // itinerary draw: :515-556
    { double r = rnd.myRand() / MAXIMUM_RANDOM_NUMBER; double sum = 0; const Dest& D = e.dests[d]; v->route = D.routes.back().route;
      for (auto& q : D.routes) { sum += q.coeff; if (r <= sum) { v->route = q.route; break; } } }
    v->route_pos = 0;
    v->lane[0] = T.first_lane + nv; v->link[0] = e.link;                                    // SetVoie / SetTuyau :688-691
    v->inst_creation = inst_creat; v->heure_entree = inst - dt + inst_creat;                // :722-724
    v->vit_reg_tuy_act = T.vit_reg;
    rnd.myRand();                                                                            // TirageJorge vehicule.cpp:747-759 (unconditional draw)
    n_created++;
    // ready / waiting lists :780-817 (CanInsertVehicle is always true for an Entree: SymuViaTripNode.h:217)
    if (e.pret[nv] != -1 || !e.attente[nv].empty()) e.attente[nv].push_back(v);
    else { e.pret[nv] = v->id; add_vehicule(v); }
    return v;
  }
  // SymCreateVehicle: Reseau::CreateVehicle reseau.cpp:13351 queues, SimuTrafic :2233-2240 activates before
  // InitVehicules: GenerateVehicle with type / lane / destination imposed (only itinerary, law and Jorge draws)
  void activate_api_vehicles(double inst) {
    for (auto& a : api_queue) {
      Entry& e = entries[a.entry]; const Link& T = links[e.link];
      int nv = a.lane;
      if (nv < 0) { double r = rnd.myRand() / MAXIMUM_RANDOM_NUMBER; double sum = 0;
        for (nv = 0; nv < T.n_lanes; nv++) { double cf = e.lane_coefs[nv]; if (cf <= 0) continue; sum += cf; if (sum >= r) break; }
        if (nv >= T.n_lanes) nv = T.n_lanes - 1; }
      Veh* v = new_vehicle(a.cls, a.id);
      { double r = rnd.myRand() / MAXIMUM_RANDOM_NUMBER; double sum = 0; const Dest& D = e.dests[a.dest]; v->route = D.routes.back().route;
        for (auto& q : D.routes) { sum += q.coeff; if (r <= sum) { v->route = q.route; break; } } }
      v->route_pos = 0; v->lane[0] = T.first_lane + nv; v->link[0] = e.link;
      v->inst_creation = a.dbt; v->heure_entree = inst - dt + a.dbt; v->vit_reg_tuy_act = T.vit_reg;
      rnd.myRand(); n_created++;
      if (e.pret[nv] != -1 || !e.attente[nv].empty()) e.attente[nv].push_back(v);
      else { e.pret[nv] = v->id; add_vehicule(v); }
    }
    api_queue.clear();
  }
  void creation_vehicules(Entry& e, double inst) {                                          // :212-325
    for (;;) {
      bool creation; double inst_creat = 0;
      if (e.exponential) {
        creation = e.crit <= inst;
        if (creation) { inst_creat = e.crit - (inst - dt); if (inst - dt + inst_creat >= 0) update_crit_creation(e, inst - dt + inst_creat, true); }
      } else {
        creation = e.crit >= 1;
        if (creation) { double q = demande_value(e, inst); if (q != 0.0) { inst_creat = dt - (e.crit - 1) * dt / (q * dt); e.crit -= 1; } }
      }
      if (!creation) break;
      generate_vehicle(e, inst, inst_creat);
    }
  }
  Veh* from_id(int id) { for (Veh* v : lst) if (v->id == id) return v; return nullptr; }
  void insertion_veh_en_attente(Entry& e, double inst) {                                    // :1642-1851
    int nl = links[e.link].n_lanes;
    for (int nv = 0; nv < nl; nv++) {
      if (e.pret[nv] != -1) {
        Veh* p = from_id(e.pret[nv]);
        if (!p) continue;
        if (p->pos[0] > p->generation_pos) {
          e.pret[nv] = -1;
          double ie = inst;
          if (p->vit[0] > 0) { ie -= dst_parcourue_ex(p) / p->vit[0]; ie = std::max<double>(p->heure_entree, ie); }
          e.last_inst_creation = ie;
        } else continue;
      }
      if (!e.attente[nv].empty()) {
        Veh* p = e.attente[nv].front();
        // CanInsertVehicle() == true -> first branch :1747-1789
        std::vector<int> ids; calcul_trafic_ex(p, inst, ids);
        if (p->pos[0] > p->generation_pos) {
          e.attente[nv].pop_front(); add_vehicule(p);
          double ie = inst;
          if (p->vit[0] > 0) { ie -= dst_parcourue_ex(p) / p->vit[0]; ie = std::max<double>(p->heure_entree, ie); }
          e.last_inst_creation = ie;
        } else { p->pos[0] = UNDEF_POSITION; p->vit[0] = p->vit_max; p->inst_creation = 0; p->deja_calcule = false; }
      }
    }
  }
  double dst_parcourue_ex(Veh* v) {                                                         // vehicule.cpp:3879-3915
    if (v->link[0] == v->link[1]) return v->pos[0] - v->pos[1];
    if (v->link[1] < 0 && v->lane[1] == v->lane[0]) return v->pos[0];
    double d;
    if (v->lane[1] >= 0) d = lanes[v->lane[1]].length - v->pos[1]; else return v->vit[0] * dt;
    for (int u : v->used_lanes) if (v->lane[0] != u) d += lanes[u].length;
    return d + v->pos[0];
  }

  // ------------------------------------------------------------------------------------------
  // Vehicule::FinCalculTrafic vehicule.cpp:4254-4583
  void fin_calcul_trafic(Veh* v, double inst) {
    size_t T = cls.size();
    if (v->link[1] != v->link[0] || v->link[0] == v->next_link) {                          // :4309-4325
      if (v->link[0] >= 0 && links[v->link[0]].brick >= 0) {
        // keep next_link only if it is a downstream link of the brick
        bool aval = v->next_link >= 0 && links[v->next_link].up_node == links[v->link[0]].brick;
        if (!aval) v->next_link = -1;
      } else v->next_link = -1;
    }
    if (v->link[1] < 0 || v->lane[1] < 0) v->dst_parcourue = v->pos[0];                     // :4328-4333
    else if (v->link[0] < 0) v->dst_parcourue += lanes[v->lane[1]].length - v->pos[1];
    else v->dst_parcourue += v->vit[0] * dt;
    if ((v->link[1] >= 0 || v->link[0] >= 0) && v->link[1] != v->link[0]) {                // :4337-4442 ('iti' counters)
      int tp = v->link[1], tn = v->link[0];
      if (tp >= 0 && links[tp].brick < 0) nb_sorti[tp * T + v->cls]++;
      if (tn >= 0 && links[tn].brick < 0) nb_entre[tn * T + v->cls]++;
    }
    // route cursor: index of the last external link entered (GetItineraireIndex on a repeat-free route)
    if (v->link[0] >= 0 && links[v->link[0]].brick < 0 && v->link[0] != v->link[1]) {
      int k = route_index(v, v->link[0]); if (k >= 0) v->route_pos = k;
    }
    if (!v->b_deceleration) v->deceleration = -1;                                           // :4555-4559
    if (v->end_leg) {                                                                        // :4577-4581 -> TripNode::VehiculeEnter TripNode.cpp:167-177
      int dl = v->dest_enter_lane;
      v->link[0] = -1; v->lane[0] = -1; v->deja_calcule = true; v->inst_sortie = v->inst_dest_reached;
      // Sortie::VehiculeEnter sortie.cpp:119-130
      for (auto& x : exits) if (x.link == lanes[dl].link) {
        lanes[dl].next_inst_sortie = next_enter_instant(x, links[lanes[dl].link].n_lanes, v->inst_dest_reached, inst, dt);
        break; }
      v->regime_fluide = true;
    }
  }
  // Reseau::SupprimeVehicules reseau.cpp:3984-4216 (order preserved)
  void supprime_vehicules() {
    exited.clear();
    std::vector<Veh*> keep; keep.reserve(lst.size());
    for (Veh* v : lst) { if (v->link[0] < 0) { exited.push_back({v->id, v->inst_sortie}); v->in_list = false; } else keep.push_back(v); }
    // Storage of a vehicle removed at step k is released at the end of step k+1: until every survivor has
    // rebuilt its context, Ctx::leader may still point at it (the reference holds shared_ptrs).
    if (!graveyard.empty()) {
      owned.erase(std::remove_if(owned.begin(), owned.end(), [](const std::unique_ptr<Veh>& p) { return p->dead == 2; }), owned.end());
      graveyard.clear();
    }
    if (keep.size() != lst.size()) {
      for (Veh* v : lst) if (v->link[0] < 0) { v->dead = 2; graveyard.push_back(v); }
      lst.swap(keep);
    }
  }


  // ==========================================================================================
  // Merge points (row A10)
  // convergent.cpp:77-100, :310-329, :333-990, :1087-1403 ; carFollowing/AbstractCarFollowing.cpp:336-664 ;
  // CarrefourAFeuxEx.cpp:1386-1593 ; reseau.cpp:2432-2600, :9614-9690 ; sensors/SensorsManager.cpp:462-700 ;
  // vehicule.cpp:4060-4160 (IsItineraire), :4615-4660 (CalculPosition)
  // ==========================================================================================
  static double min_d(double a, double b) { return (b < a) ? b : a; }                       // std::min<double> (matters when an operand is NaN)
  double cvg_tf(const Cvg& c, int k) const { return std::max<double>(c.tf[k], 1 / cls[k].debit_max()); }   // Convergent::GetTf :310-329 (every type is listed)
  int pt_of_upstream_lane(const Cvg& c, int lane) const {                                  // GetPointDeConvergence(VoieMicro*) :1340-1353
    for (int p : c.pts) for (auto& a : pts[p].vam) if (a.lane == lane) return p;
    return -1; }
  bool voie_amont_non_prioritaire(const Cvg& c, int lane) const {                          // :1230-1258
    int p = pt_of_upstream_lane(c, lane); int niv = 0, nmin = INT32_MAX;
    if (p >= 0) for (auto& a : pts[p].vam) { if (a.lane == lane) niv = a.niv; if (nmin > a.niv) nmin = a.niv; }
    return niv > nmin && niv > 1; }
  double inst_last_passage(const Cvg& c, int down_lane) const {                            // :1170-1184 (UNDEF_INSTANT when the lane is not a point of this merge)
    for (int p : c.pts) if (pts[p].down_lane == down_lane) return pts[p].inst_last_passage;
    return -DBL_MIN; }
  double calcul_vitesse_leader(const Cvg& c, Veh* L, double inst, double pas) {            // :1121-1168
    if (L->next_voie < 0) L->next_voie = calcul_next_voie(L, L->lane[1]);
    double tl = inst_last_passage(c, L->next_voie);
    double t0 = tl >= 0 ? tl + cvg_tf(c, L->cls) - (inst - pas) : 0;
    double p1 = L->pos[1] <= UNDEF_POSITION ? 0 : L->pos[1];
    if (t0 > 0) return (lanes[L->lane[1]].length - p1) / t0;
    return L->vit_max; }
  Veh* vehicule_en_attente(const Cvg& c, int lane) const {                                 // GetVehiculeEnAttente :1087-1116
    for (Veh* v : c.ins_veh) { if (v->lane[1] == lane && v->lane[0] != lane) return v;
      for (int u : v->used_lanes) if (u == lane && v->lane[0] != lane) return v; }
    return nullptr; }
  // a vehicle that does not insert stays at the end of its upstream lane (:405-412, :433-438, :635-644, :819-827, :972-987)
  void immobilise(Veh* v, int lane, bool acc_sic) {
    v->pos[0] = lanes[lane].length - 0.0001; v->link[0] = lanes[lane].link; v->lane[0] = lane;
    v->vit[0] = dst_parcourue_ex(v) / dt;
    set_acc(v, acc_sic ? v->vit[1] / dt : (v->vit[0] - v->vit[1]) / dt);
    update_used_lanes(v); n_pins++; }
  // sensors of a merge: SensorsManager::GetDebitMoyen / GetNbVehTotal / GetRatioTypeVehicule (:551-700)
  int capteur_of(const Cvg& c, int link) const { if (link == c.down_link) return 0;       // Convergent::GetCapteur :1000-1014
    for (size_t k = 0; k < c.cpt.size(); k++) if (c.cpt[k].link == link) return (int)k;
    return -1; }
  double cpt_somme(const Cvg& c, int k, int nvoie, int type /* -1 = all */) const {
    const Capteur& q = c.cpt[k]; int nl = links[q.link].n_lanes; if (nvoie < 0 || nvoie >= nl) return 0;
    int kmax = std::min(c.nb_inst, c.window); double s = 0;
    for (size_t t = 0; t < cls.size(); t++) if (type < 0 || (int)t == type) for (int j = 0; j < kmax; j++) s += q.cnt[(t * nl + nvoie) * c.window + j];
    return s; }
  double debit_moyen(const Cvg& c, int k, int nvoie) const { if (k < 0) return 0; return cpt_somme(c, k, nvoie, -1) / (std::min(c.nb_inst, c.window) * dt); }
  double nb_veh_total(const Cvg& c, int k, int nvoie) const { if (k < 0) return 0; return cpt_somme(c, k, nvoie, -1); }
  double ratio_type(const Cvg& c, int k, int nvoie, int type) const { if (k < 0) return 0; double tot = cpt_somme(c, k, nvoie, -1);
    return tot > 0 ? cpt_somme(c, k, nvoie, type) / tot : 0; }
  // Vehicule::IsItineraire vehicule.cpp:4060-4160 ('iti' mode; itineraries hold external links only)
  bool is_itineraire(const Veh* v, int link) const {
    const auto& r = routes[v->route];
    for (int k : r) if (k == link) return true;
    int br = links[link].brick;
    if (br >= 0) for (size_t i = 0; i + 1 < r.size(); i++) if (links[r[i]].down_node == br) {
      for (int mi : node_mvts[br]) { const CafMvt& m = caf_mvts[mi]; if (m.ent_link == r[i] && m.exit_link == r[i + 1]) for (int t : m.tl) if (t == link) return true; } }
    return false; }
  // Reseau::GetVehiculeAmont reseau.cpp:9638-9690 (first lane of every link only)
  Veh* get_vehicule_amont(int link, double dst, double cum) {
    Veh* veh = get_near_amont_vehicule(links[link].first_lane, links[link].length, -1.0, nullptr);
    if (veh) return veh;
    if (dst < cum) return nullptr;
    if (node_type[links[link].up_node] == 'E' || node_type[links[link].up_node] == 'S' || node_type[links[link].up_node] == 'P') return nullptr;
    for (int am : lcam[link]) {
      veh = get_vehicule_amont(am, dst, cum + links[link].length);
      if (veh && dst > cum + links[link].length - veh->pos[1] && calcul_next_voie(veh, veh->lane[1]) == links[link].first_lane) return veh; }
    return nullptr; }
  double max_rap_vit_sur_qx(int link) const { return std::min(links[link].vit_reg, cls[0].vx); }   // TuyauMicro::GetMaxRapVitSurQx tuyau.cpp:1944-1958
  // Reseau::GetPrevVehicule(pVeh, bSearchNextVoie = true) reseau.cpp:3650-3820
  Veh* get_prev_vehicule_next(Veh* v) {
    int lane = v->lane[1]; if (lane < 0) return nullptr;
    Veh* r = get_prev_vehicule(v); if (r) return r;
    std::vector<Veh*> l; int nvoie = lanes[lane].num; int next_link = -1; int voie = lane;
    char ta = links[v->link[1]].type_aval;
    if (ta != 'S' && ta != 'P') next_link = v->link[1];
    auto scan = [&](int ln) { double pv = lanes[ln].length + 1;
      for (Veh* c : lane_set[ln]) if (c->pos[1] <= pv && c->pos[1] >= 0 && c != v) { if (c->pos[1] == pv) l.push_back(c); else { l.resize(1); l[0] = c; pv = c->pos[1]; } } };
    if (ta == 'C' && !lcav[v->link[1]].empty()) {                                          // :3726-3768
      next_link = lcav[v->link[1]].front();
      voie = calcul_next_voie(v, voie);
      if (voie < 0) voie = links[next_link].first_lane + std::min(nvoie, links[next_link].n_lanes - 1);
      scan(voie);
    }
    if (l.empty() && next_link >= 0) for (int nn : lcav[next_link]) if (is_itineraire(v, nn)) {   // :3771-3815
      voie = calcul_next_voie(v, voie);
      if (voie < 0) voie = links[nn].first_lane + std::min(nvoie, links[nn].n_lanes - 1);
      scan(voie);
    }
    return get_last_vehicule(l); }
  // Vehicule::CalculPosition vehicule.cpp:4615-4760
  double calcul_position(const Veh* v, double tr, int lane) const {
    double p = v->pos[1] + v->vit[0] * tr;
    if (lanes[lane].link == lanes[v->lane[1]].link) return p;
    std::vector<int> iti; const auto& r = routes[v->route];                                // Reseau::ComputeItineraireComplet reseau.cpp:14002-14032
    for (size_t i = 0; i < r.size(); i++) { if (i > 0 && links[r[i - 1]].type_aval != 'S') {
        int br = links[r[i - 1]].down_node;
        if (node_type[br] == 'C' && links[r[i]].up_node == br) {                            // CarrefourAFeuxEx::GetTuyauxInternes :1598-1660: first lane pair that has a movement
          bool found = false;
          for (int a = 0; a < links[r[i - 1]].n_lanes && !found; a++) for (int b2 = 0; b2 < links[r[i]].n_lanes && !found; b2++)
            for (int mi : node_mvts[br]) { const CafMvt& m = caf_mvts[mi]; if (m.ent_link == r[i - 1] && m.exit_link == r[i] && !m.tl.empty() &&
                lanes[links[m.tl.front()].first_lane].prev_lane == links[r[i - 1]].first_lane + a) {
              int last = links[m.tl.back()].first_lane; bool to_b = false; for (auto& q : lanes[last].succ) if (q.next_lane == links[r[i]].first_lane + b2) to_b = true;
              if (to_b) { for (int t : m.tl) iti.push_back(t); found = true; break; } } } } }
      iti.push_back(r[i]); }
    int idx = -1; for (size_t k = 0; k < iti.size(); k++) if (iti[k] == lanes[v->lane[1]].link) { idx = (int)k; break; }
    if (idx != -1) { p = -(lanes[v->lane[1]].length - p);
      for (size_t j = idx + 1; j < iti.size(); j++) { if (iti[j] == lanes[lane].link) return p; p -= links[iti[j]].length; } }
    return 9999; }

  // AbstractCarFollowing::TestConvergentInsertion carFollowing/AbstractCarFollowing.cpp:336-664 (no roundabout, calcul_tq_convergent off)
  bool test_convergent_insertion(Veh* v, PtCvg& P, Cvg& C, int TAmP, int VAmP, int VAmNP, int nVoieGir, double inst, int j,
                                 double demande1, double demande2, double offre, Veh* leader, Veh* follower, double& tr, double& ta) {
    bool res = false;
    if (!P.free_flow) {                                                                     // :343-377
      double q2g = C.gamma * offre / (1 + C.gamma);
      if (demande2 < q2g) { res = true; P.rand_numbers = 0; }
      else { double proba = q2g * dt; res = false; n_cong++;
        rnd.myRand();                                                                       // one draw, unused
        for (int k = 0; k < P.rand_numbers; k++) { double r = rnd.myRand() / MAXIMUM_RANDOM_NUMBER; P.rand_numbers = P.rand_numbers - 1;
          if (r < proba) { res = true; break; } } }
      return res;
    }
    P.rand_numbers = 0;                                                                     // :381
    const Cls& k = cls[v->cls];
    const double tmin = std::max<double>(cvg_tf(C, v->cls), 1 / k.debit_max());
    if (!(inst - P.inst_last_passage >= tmin)) return false;                                 // :385-391 downstream condition
    const int TAv = C.down_link;
    if (follower) {                                                                          // :395-461
      if (node_type[links[TAmP].up_node] == 'C' && links[TAmP].brick >= 0) {                 // get_Type_amont() == 'D': internal link behind a CAF divergent
        int ent = lcam[TAmP].empty() ? -1 : lcam[TAmP].front();
        if (ent >= 0 && lcav[ent].size() > 1) {                                              // m_LstVoiesSvt only ever holds results of CalculNextVoie: no entry lane, the search fails
          calcul_next_voie(follower, links[ent].first_lane);                                 // :432 (may consume the memoised draw of that lane)
        }
      } else if (!is_itineraire(follower, TAv)) { follower = nullptr; res = true; }         // :455-459
    }
    Veh* foll_int = nullptr;                                                                 // :472-500
    if (links[TAmP].n_lanes > 1 && links[TAv].n_lanes > 1 && nVoieGir == 0) {
      foll_int = get_near_amont_vehicule(links[TAmP].first_lane + nVoieGir + 1, lanes[VAmP].length, -1.0, nullptr);
      if (foll_int) { if (foll_int->res_tir_foll_int == -1) foll_int->res_tir_foll_int = (double)rnd.myRand() / MAXIMUM_RANDOM_NUMBER;
        foll_int->res_tir_foll_int = -1; foll_int = nullptr; }                               // not a roundabout: never in the way
    }
    const int TNP = lanes[VAmNP].link;
    double tfbar = 0;                                                                        // :505-513
    if (nb_veh_total(C, capteur_of(C, TNP), j) == 0) tfbar = cvg_tf(C, v->cls);
    else for (size_t i = 0; i < cls.size(); i++) tfbar += ratio_type(C, capteur_of(C, TNP), j, (int)i) * cvg_tf(C, (int)i);
    double tmbar = 0;
    for (size_t i = 0; i < cls.size(); i++) tmbar += ratio_type(C, capteur_of(C, TAmP), 0, (int)i) * (1 / cls[i].debit_max());
    double q1mu = 1 / ((tfbar * C.mu) + tmbar);
    double vit_am;                                                                           // :520-525
    { Veh* w = get_vehicule_amont(TAmP, max_rap_vit_sur_qx(TAmP), 0);
      if (w && w->link[1] >= 0) vit_am = std::min<double>(w->vit_max, vit_reg(w->link[1]));
      else vit_am = std::min<double>(max_vit_max, vit_reg(TAmP)); }
    double dlag;
    if (demande1 <= q1mu) dlag = (-k.w + vit_am) / (-k.kx * k.w);
    else { double eta = (demande1 - q1mu) / (1 / tmbar - q1mu); dlag = (-k.w + vit_am) / (-k.kx * k.w) * (1 - eta); }   // :537-538 (the first assignment is dead)
    if (VAmNP == v->lane[1]) tr = dt * (lanes[v->lane[1]].length - v->pos[1]) / dst_parcourue_ex(v);            // :544-547
    else tr = dt * (lanes[v->lane[1]].length - v->pos[1] + lanes[VAmNP].length) / dst_parcourue_ex(v);
    const double plim = lanes[VAmP].length - dlag;
    if (inst - dt + tr > P.inst_last_passage + tmin) {                                       // :550-587
      if (follower || foll_int) {
        double pf = 9999; if (follower) pf = calcul_position(follower, tr, VAmP);
        if (pf <= plim) { res = true; ta = dt - tr; } else res = false;
      } else { res = true; ta = dt - tr; }
    } else {                                                                                 // :589-628
      tr = dt - (inst - (P.inst_last_passage + tmin));
      if (!follower && !foll_int) { ta = dt - tr; res = true; }
      else { double pf = 9999; if (follower) pf = calcul_position(follower, tr, VAmP);
        if (pf <= plim) { ta = dt - tr; res = true; } else res = false; }
    }
    return res;
  }

  // PointDeConvergence::CalculInsertion convergent.cpp:333-990
  void calcul_insertion(int pi, double inst) {
    PtCvg& P = pts[pi]; Cvg& C = cvgs[P.cvg];
    P.rand_numbers++;                                                                       // :366
    if (P.vam.empty()) return;
    int VAmP = -1; { int nt = INT32_MAX; for (auto& a : P.vam) if (a.niv < nt) { VAmP = a.lane; nt = a.niv; } }   // GetVoieAmontPrioritaire :77-100
    int niv_prio = -1; for (auto& a : P.vam) if (a.lane == VAmP) { niv_prio = a.niv; break; }
    int niv = 999; Veh* att = nullptr; int VAmNP = -1; int att_lane = -1; bool have = false;   // LstVehEnAttente holds at most one entry (the Tmp list is cleared per lane, :447)
    for (size_t i = 0; i < P.vam.size(); i++) if (P.vam[i].niv > niv_prio) {                // :376-449
      int VAm = P.vam[i].lane;
      Veh* w = vehicule_en_attente(C, VAm);
      if (!(w && w->vit[0] > 0)) w = nullptr;
      if (w) {
        if (P.vam[i].niv < niv) {
          if (have && att) immobilise(att, att_lane, true);                                 // :399-413 (acc = vit[1]/dt, sic)
          att = w; att_lane = VAm; have = true; VAmNP = VAm; niv = P.vam[i].niv;
        } else immobilise(w, VAm, true);                                                    // :425-442
      }
    }
    if (!att) return;                                                                        // :452-460
    Veh* follower = get_near_amont_vehicule(VAmP, lanes[VAmP].length + 0.01, dt * max_vit_max, nullptr);   // :463
    int TAmP = lanes[VAmP].link; int nVoieGir = lanes[VAmP].num;
    auto drop = [&](Veh*& f) { if (f) { if (f->link[1] < 0) f = nullptr; else if (f->arret_au_feu) f = nullptr; else if (!f->ctx) f = nullptr; } };
    drop(follower);                                                                          // :469-483
    if (!follower) for (size_t i = 0; i < P.vam.size(); i++) if (P.vam[i].niv < niv && P.vam[i].niv > 1) {   // :486-504
      follower = get_near_amont_vehicule(P.vam[i].lane, lanes[P.vam[i].lane].length, -1.0, nullptr);
      if (follower && !follower->arret_au_feu) { niv = P.vam[i].niv; VAmP = P.vam[i].lane; TAmP = lanes[VAmP].link; nVoieGir = lanes[VAmP].num; }
      else follower = nullptr; }
    drop(follower);                                                                          // :507-521
    if (follower && !is_itineraire(follower, TAmP)) follower = nullptr;                     // :525-528
    Veh* leader = get_near_aval_vehicule(P.down_lane, 0, nullptr);                          // :531-533
    if (!leader && att->link[1] >= 0) leader = get_prev_vehicule_next(att);
    { Veh* w = get_near_amont_vehicule(P.down_lane, C.pos_cpt_av, -1.0, nullptr);           // :538-550
      if (w && w->lane[1] != P.down_lane) w = nullptr;
      P.free_flow = w ? w->regime_fluide : true; }
    if (!P.free_flow && !get_vehicule_amont(TAmP, max_rap_vit_sur_qx(TAmP), 0)) P.free_flow = true;   // :554-556
    double offre = 0, demande1 = 0, demande2 = 0;
    if (!P.free_flow) {                                                                      // :560-583
      offre = min_d(max_debit_max, debit_moyen(C, capteur_of(C, C.down_link), nVoieGir));
      int TNP = lanes[VAmNP].link; bool np_free;
      Veh* w = get_vehicule_amont(TNP, max_rap_vit_sur_qx(TNP), 0);
      if (w) np_free = !(w->vit[1] < std::min<double>(w->vit_max, vit_reg(TNP))); else np_free = true;
      if (np_free) demande2 = min_d(max_debit_max, debit_moyen(C, capteur_of(C, TNP), std::min<int>(lanes[VAmNP].num, nVoieGir)));
      else demande2 = max_debit_max;
    } else demande1 = min_d(max_debit_max, debit_moyen(C, capteur_of(C, TAmP), nVoieGir));  // :586
    double tr = 0, ta = 0;
    bool ins = test_convergent_insertion(att, P, C, TAmP, VAmP, VAmNP, nVoieGir, inst, 0, demande1, demande2, offre, leader, follower, tr, ta);   // :590-600
    // TestConvergentInsertion drops its own copy of the follower only (pVehFollower is passed by value)
    P.insertion = ins;
    if (trace_merge) fprintf(stderr, "merge t=%g pt %d down %d: att %d (lane %d niv %d) prio lane %d (niv %d) follower %d leader %d free %d ins %d tr %g ta %g d1 %g\n", inst, pi, P.down_lane, att->id, VAmNP, niv, VAmP, niv_prio, follower ? follower->id : -1, leader ? leader->id : -1, (int)P.free_flow, (int)ins, tr, ta, demande1);
    if (!ins) { immobilise(att, VAmNP, false); n_refus++; return; }                          // :967-987
    if (P.free_flow) {                                                                       // :656-836
      Ctx ic; ic.post = true; ic.start_pos = 0;
      if (leader && leader->link[1] >= 0) { double pl = leader->pos[1] + (dt - ta) * leader->vit[0];
        if (P.down_lane != leader->lane[1]) pl += lanes[P.down_lane].length;
        ic.leader = leader; ic.leader_id = leader->id; ic.leader_dist = pl; }
      { bool add = false; for (int ln : att->ctx->lanes) { if (lanes[ln].link == lanes[P.down_lane].link) add = true; if (add) ic.lanes.push_back(ln); } }
      double pos_att = ic.lanes.empty() ? 0.0 : compute_law(att, ta, ic);
      if (pos_att > 0) {
        double dfol = 0; 
        if (follower) {                                                                      // :704-749
          Ctx fc; fc.post = true;
          double ptr = follower->pos[1] + (dt - ta) * follower->vit[0];
          std::vector<int> lf = follower->ctx->lanes;
          while (lf.size() > 1 && ptr > links[lanes[lf.front()].link].length) { ptr -= links[lanes[lf.front()].link].length; lf.erase(lf.begin()); }
          double dconv = 0;
          for (size_t i = 0; i < lf.size(); i++) { if (i == 0) dconv = lanes[lf[i]].length - ptr; else dconv += lanes[lf[i]].length; if (lanes[lf[i]].link == TAmP) break; }
          fc.leader = att; fc.leader_id = att->id; fc.leader_dist = dconv; fc.lanes = lf; fc.start_pos = ptr;
          fc.deltaN = follower->ctx->deltaN; fc.contraction = follower->ctx->contraction;
          dfol = compute_law(follower, ta, fc);
          dfol = std::min<double>(dst_parcourue_ex(follower), (dt - ta) * follower->vit[0] + dfol);
        }
        if (!follower || dfol >= 0) {
          bool dont_move = false; double trav = 0;                                          // :753-786
          for (size_t i = 0; i < ic.lanes.size(); i++) { int ln = ic.lanes[i]; double llen = len_ex(ln); double ep = pos_att - trav;
            if (ln == att->lane[0] && ep > att->pos[0]) { dont_move = true; break; }
            else if (ep <= llen) { att->link[0] = lanes[ln].link; att->lane[0] = ln; att->pos[0] = ep; update_used_lanes(att); break; }
            trav += std::min<double>(ep, llen); }
          if (!dont_move) { att->vit[0] = (pos_att + (dt - ta) * att->vit[0]) / dt; set_acc(att, (att->vit[0] - att->vit[1]) / dt); }
          if (follower) {                                                                    // :788-816
            trav = 0; const std::vector<int>& lf2 = follower->ctx->lanes;
            for (size_t i = 0; i < lf2.size(); i++) { int ln = lf2[i]; double sp = i == 0 ? follower->pos[1] : 0; double llen = len_ex(ln); double ep = dfol - trav + sp;
              if (ep <= llen) { follower->link[0] = lanes[ln].link; follower->lane[0] = ln; follower->pos[0] = ep; update_used_lanes(follower); break; }
              trav += std::min<double>(ep, llen) - sp; }
            follower->vit[0] = (dfol + (dt - ta) * follower->vit[0]) / dt; set_acc(follower, (follower->vit[0] - follower->vit[1]) / dt);
          }
          P.inst_last_passage = inst - dt + tr; n_insertions++;
          return;
        }
      }
      immobilise(att, VAmNP, false);                                                         // :819-830
      for (auto it = C.ins_veh.begin(); it != C.ins_veh.end(); ++it) if (*it == att) { C.ins_veh.erase(it); break; }   // DelInsVeh
      return;
    }
    // congested :838-966
    if (!leader) return;                                                                     // the reference dereferences the leader here; a congested downstream lane always has one
    double dconv = att->link[1] == lanes[VAmNP].link ? lanes[VAmNP].length - att->pos[1]
                                                     : lanes[VAmNP].length + lanes[att->lane[1]].length - att->pos[1];
    Ctx ic; ic.post = true;
    double pl = leader->pos[1]; if (P.down_lane != leader->lane[1]) pl += lanes[P.down_lane].length;
    pl += dconv;
    ic.leader = leader; ic.leader_id = leader->id; ic.leader_dist = pl; ic.lanes = att->ctx->lanes; ic.start_pos = att->pos[1];   // deltaN = 1 (ComputeDeltaN without a previous context)
    double d = compute_law(att, dt, ic);
    d = std::min<double>(dst_parcourue_ex(att), d);
    if (!(d >= dconv)) return;                                                               // :962-965 "PAS NORMAL"
    { double trav = 0; const std::vector<int>& ll = ic.lanes;                               // :880-899
      for (size_t i = 0; i < ll.size(); i++) { int ln = ll[i]; double sp = i == 0 ? att->pos[1] : 0; double llen = len_ex(ln); double ep = d - trav + sp;
        if (ep <= llen) { att->link[0] = lanes[ln].link; att->lane[0] = ln; att->pos[0] = ep; update_used_lanes(att); break; }
        trav += std::min<double>(ep, llen) - sp; }
      att->vit[0] = d / dt; set_acc(att, (att->vit[0] - att->vit[1]) / dt); }
    P.inst_last_passage = inst - dt; n_insertions++;
    if (follower) {                                                                          // :908-958
      double pos = follower->link[1] == TAmP ? lanes[VAmP].length - follower->pos[1]
                                             : lanes[VAmP].length + lanes[follower->lane[1]].length - follower->pos[1];
      Ctx fc; fc.post = true; fc.leader = att; fc.leader_id = att->id; fc.leader_dist = pos; fc.lanes = follower->ctx->lanes; fc.start_pos = follower->pos[1];
      double dfol = std::max<double>(0, compute_law(follower, dt, fc));
      if (dst_parcourue_ex(follower) > dfol) {
        double trav = 0;
        for (size_t i = 0; i < fc.lanes.size(); i++) { int ln = fc.lanes[i]; double sp = i == 0 ? follower->pos[1] : 0; double llen = len_ex(ln); double ep = dfol - trav + sp;
          if (ep <= llen) { follower->link[0] = lanes[ln].link; follower->lane[0] = ln; follower->pos[0] = ep; update_used_lanes(follower); break; }
          trav += std::min<double>(ep, llen) - sp; }
        follower->vit[0] = dfol / dt; set_acc(follower, (follower->vit[0] - follower->vit[1]) / dt);
      }
    }
  }

  // CarrefourAFeuxEx::UpdateConvergents CarrefourAFeuxEx.cpp:1386-1593 (start of the step, reseau.cpp:2225-2229)
  void caf_update_convergents(double inst) {
    if (cvgs.empty()) return;
    std::vector<char> occ(lanes.size(), 0);                                                // Reseau::HasVehicule :9614-9635: lane at the end of the previous step
    for (Veh* v : lst) if (v->lane[0] >= 0) occ[v->lane[0]] = 1;
    for (auto& C : cvgs) { if (C.node < 0 || node_ctrl[C.node] < 0) continue;               // :1406-1407 no controller: priorities stay
      for (int p : C.pts) for (auto& a : pts[p].vam) a.niv = a.niv_ref; }
    for (auto& C : cvgs) { if (C.node < 0 || node_ctrl[C.node] < 0) continue;
      for (int p : C.pts) { PtCvg& P = pts[p];
        if (P.vam.size() == 0) break;
        if (P.vam.size() == 1) { P.vam[0].niv = 1; break; }
        std::vector<const CafMvt*> g1, g2, g3; std::vector<int> tmv;                        // :1437-1497
        for (int mi : node_mvts[C.node]) { const CafMvt& m = caf_mvts[mi];
          int tm = -1; for (int t : m.tl) { for (auto& a : P.vam) if (lanes[a.lane].link == t) { tm = t; break; } if (tm >= 0) break; }
          if (tm < 0) continue;
          double pr = 0;
          if (statut_couleur(node_ctrl[C.node], inst, m.ent_link, m.exit_link, &pr) == 1) g2.push_back(&m);
          else { bool veh = false; for (int t : m.tl) { if (occ[links[t].first_lane]) { veh = true; break; } if (t == tm) break; }
            if (veh) g1.push_back(&m); else g3.push_back(&m); } }
        const int n = (int)P.vam.size();
        auto bump = [&](const CafMvt* m, int add) { for (int t : m->tl) for (auto& a : P.vam) if (lanes[a.lane].link == t) { a.niv = a.niv_ref + add; break; } };
        if (g1.empty()) { for (auto* m : g3) bump(m, n + 1); }                               // :1499-1525
        else { for (auto* m : g2) bump(m, n + 1); for (auto* m : g3) bump(m, 2 * n + 1); }   // :1527-1588
      } }
  }
  // SensorsManager::UpdateTabNbVeh sensors/SensorsManager.cpp:462-527 (Convergent::UpdateCapteurs, reseau.cpp:14425-14428)
  void update_tab_nb_veh(Cvg& C) {
    int last = C.window - 1;
    if (C.nb_inst < last) { C.nb_inst++; last = C.nb_inst; }
    else for (auto& q : C.cpt) { size_t rows = q.cnt.size() / (size_t)std::max(1, C.window);
      for (size_t r = 0; r < rows; r++) { double* a = &q.cnt[r * C.window]; for (int j = 0; j + 1 < C.window; j++) a[j] = a[j + 1]; if (C.window > 0) a[C.window - 1] = 0; } }
    C.last_ind = last; }
  // Reseau::UpdateConvergents reseau.cpp:2432-2600 (offre_aval_convergent_deltaN off: integer counts)
  void update_convergents_capteurs() {
    const size_t T = cls.size();
    auto hit = [&](const std::pair<int, int>& k, int cl, int lane) { Cvg& C = cvgs[k.first]; Capteur& q = C.cpt[k.second]; int nl = links[q.link].n_lanes;
      if (C.last_ind >= 0 && C.last_ind < C.window) q.cnt[((size_t)cl * nl + lanes[lane].num) * C.window + C.last_ind] += 1; (void)T; };
    for (Veh* v : lst) { double p0 = v->pos[0];
      if (!(v->vit[0] > 0 && p0 >= 0)) continue;
      int t0 = v->link[0], t1 = v->link[1]; double p1 = v->pos[1];
      if (t1 >= 0) for (auto& k : link_capteurs[t1]) { double x = cvgs[k.first].cpt[k.second].pos;
        if (p1 < x && (p0 >= x || t1 != t0) && v->lane[1] >= 0) hit(k, v->cls, v->lane[1]); }
      if (t1 != t0) {
        if (t0 >= 0) for (auto& k : link_capteurs[t0]) { if (p0 >= cvgs[k.first].cpt[k.second].pos && v->lane[0] >= 0) hit(k, v->cls, v->lane[0]); }
        for (int u : v->used_lanes) for (auto& k : link_capteurs[lanes[u].link]) hit(k, v->cls, u); }
    }
  }

  // ------------------------------------------------------------------------------------------
  // Reseau::SimuTrafic reseau.cpp:2028-2428 + SimuTraficMicroMacro :14152-14470
  int step() {
    max_depth = 0;
    t += dt; ninst++;                                                                        // :2105-2106
    for (auto& e : entries) update_crit_creation_step(e);                                   // :2219-2223
    caf_update_convergents(t);                                                               // :2225-2229
    activate_api_vehicles(t);                                                                // :2233-2240
    init_vehicules();                                                                        // :2250
    for (auto& e : entries) creation_vehicules(e, t);                                       // :2255 -> :14143
    if (chgtvoie) changement_de_voie();                                                      // :14181 (ComputeFirstVehicules :14184 only feeds flux nodes)
    for (size_t i = 0; i < lst.size(); i++) { std::vector<int> ids; calcul_trafic_ex(lst[i], t, ids); }   // MiseAJourVehicules :3827-3865
    for (auto& e : entries) insertion_veh_en_attente(e, t);                                 // :14293-14296
    for (auto& C : cvgs) { for (int p : C.pts) calcul_insertion(p, t); C.ins_veh.clear(); }  // :14310-14314 (Liste_convergents: std::map order)
    for (Veh* v : lst) fin_calcul_trafic(v, t);                                             // :14364 -> :3871
    supprime_vehicules();                                                                    // :14418
    for (auto& C : cvgs) update_tab_nb_veh(C);                                              // :14425-14428
    update_convergents_capteurs();                                                           // :14458
    for (Veh* v : lst) if (v->link[0] >= 0 && v->lane[0] >= 0) v->vit_reg_tuy_act = links[v->link[0]].vit_reg;   // UpdateVitessesReg :2606-2627
    return (int)lst.size();
  }
  void update_crit_creation_step(Entry& e) {
    if (e.exponential) update_crit_creation(e, t, false); else update_crit_creation(e, t, false);
  }
};

}  // namespace

// ---- C API (ctypes) -------------------------------------------------------------------------------
extern "C" {
void* orc_load(const char* path) { auto* s = new Sim(); if (!s->load(path)) { delete s; return nullptr; } return s; }
void orc_free(void* h) { delete (Sim*)h; }
int orc_step(void* h) { return ((Sim*)h)->step(); }
int orc_nveh(void* h) { return (int)((Sim*)h)->lst.size(); }
double orc_time(void* h) { return ((Sim*)h)->t; }
unsigned orc_rng_count(void* h) { return ((Sim*)h)->rnd.count; }
long orc_counter(void* h, int which) { Sim* s = (Sim*)h; return which == 0 ? s->n_ties : which == 1 ? s->n_loops : which == 2 ? s->n_created : which == 3 ? s->max_depth :
  which == 4 ? s->n_insertions : which == 5 ? s->n_refus : which == 6 ? s->n_pins : which == 7 ? s->n_cong : 0; }
// merge points: last passage instant and draw counter of every point, window sums of every sensor (test access)
int orc_merge_state(void* h, double* inst_last, int* rand_numbers, int cap) { Sim* s = (Sim*)h; int n = (int)s->pts.size();
  for (int p = 0; p < n && p < cap; p++) { if (inst_last) inst_last[p] = s->pts[p].inst_last_passage; if (rand_numbers) rand_numbers[p] = s->pts[p].rand_numbers; } return n; }
// state of m_LstVehicles in list order after the step
void orc_dump(void* h, int* id, int* lane, int* route_pos, int* flags, double* pos, double* vit, double* acc, double* deltaN) {
  Sim* s = (Sim*)h; size_t i = 0;
  for (Veh* v : s->lst) {
    id[i] = v->id; lane[i] = v->lane[0]; route_pos[i] = v->route_pos;
    flags[i] = (v->regime_fluide ? 1 : 0) | (v->arret_au_feu ? 2 : 0) | ((v->ctx && v->ctx->free_flow) ? 4 : 0) | (v->cls << 8);
    pos[i] = v->pos[0]; vit[i] = v->vit[0]; acc[i] = v->acc[0]; deltaN[i] = v->ctx ? v->ctx->deltaN : 1.0; i++;
  }
}
int orc_leaders(void* h, int* leader_id) { Sim* s = (Sim*)h; size_t i = 0; for (Veh* v : s->lst) leader_id[i++] = v->ctx ? v->ctx->leader_id : -1; return (int)i; }
void orc_link_counts(void* h, int* entre, int* sorti) { Sim* s = (Sim*)h; size_t n = s->nb_entre.size();
  memcpy(entre, s->nb_entre.data(), n * 4); memcpy(sorti, s->nb_sorti.data(), n * 4); }
int orc_exited(void* h, int* ids, double* inst, int cap) { Sim* s = (Sim*)h; int n = 0;
  for (auto& e : s->exited) { if (n < cap) { ids[n] = e.id; inst[n] = e.inst; } n++; } return n; }
int orc_create_vehicle(void* h, int cls, int entry, int dest, int lane, double dbt) { Sim* s = (Sim*)h; int id = s->last_id++;
  s->api_queue.push_back({id, cls, entry, dest, lane, dbt}); return id; }
void orc_set_partition(void* h, const int* node_part) { Sim* s = (Sim*)h; s->lane_owner.resize(s->lanes.size());
  for (size_t l = 0; l < s->lanes.size(); l++) { const Link& T = s->links[s->lanes[l].link];
    int node = T.brick >= 0 ? T.brick : ((T.type_aval == 'S' || T.type_aval == 'P') ? T.up_node : T.down_node); s->lane_owner[l] = node_part[node]; } }
int orc_waiting(void* h) { Sim* s = (Sim*)h; int n = 0; for (auto& e : s->entries) for (auto& q : e.attente) n += (int)q.size(); return n; }
// raw RNG access for the known-answer test of the generator itself
void* orc_rng_new(unsigned seed) { auto* r = new RandManager(); r->mySrand(seed); return r; }
int orc_rng_next(void* r) { return ((RandManager*)r)->myRand(); }
unsigned orc_rng_raw(void* r) { return ((RandManager*)r)->gen.next(); }
void orc_rng_free(void* r) { delete (RandManager*)r; }
}
