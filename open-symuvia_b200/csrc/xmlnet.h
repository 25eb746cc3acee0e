// xmlnet.h — ROOT_SYMUBRUIT front end: the subset of SymuVia's input document -> flat tables (flatnet.h).
//
// Replaces, for that subset, Reseau::LoadReseauXML (symuvia/src/reseau.cpp:4931-9612, Xerces / XQilla in the reference) and the
// load-time construction it triggers: CarrefourAFeuxEx::Init (CarrefourAFeuxEx.cpp:215-1196: one Divergent per entry link, one
// single-lane internal link per allowed movement, one Convergent per exit link with several movements, geometric priority
// levels), FinInit (:1205-1275: one stop line per movement), Convergent::Init / FinInit (convergent.cpp:208-255), the allowed
// movements of the distributors and merges (Connection.h m_mapMvtAutorises), fixed-time signal plans (ControleurDeFeux.cpp),
// origins with their demand, lane / type / destination repartitions and predefined routes (usage/SymuViaTripNode.cpp), exits with
// capacities (sortie.cpp), initial vehicles (reseau.cpp:11680-11773).  Host only; no third-party parser.
//
// Subset and private extensions: one SIMULATION / TRAFIC / RESEAU / SCENARIO; comportementflux="iti";
// TYPE_DE_VEHICULE@loi (newell | gipps | idm) instead of LOIS_DE_POURSUITE; RESEAU/ROUTES/ROUTE@id with TRONCONS_/TRONCON_@id
// as the predefined itineraries; INIT/TRAJS/TRAJ@route; lane end points by lateral offset (num + 0.5) * largeur_voie to the
// right of the link axis.
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "flatnet.h"

namespace symb200 {

struct XmlNode {
  std::string name; std::vector<std::pair<std::string, std::string>> attrs; std::vector<std::unique_ptr<XmlNode>> kids;
  const std::string* attr(const char* k) const { for (auto& a : attrs) if (a.first == k) return &a.second; return nullptr; }
  std::string str(const char* k, const char* def = "") const { auto* v = attr(k); return v ? *v : std::string(def); }
  double num(const char* k, double def) const { auto* v = attr(k); if (!v || v->empty()) return def; if (*v == "infini") return 1e9; return strtod(v->c_str(), nullptr); }
  bool flag(const char* k, bool def) const { auto* v = attr(k); if (!v) return def; return *v == "true" || *v == "1"; }
  const XmlNode* kid(const char* n) const { for (auto& c : kids) if (c->name == n) return c.get(); return nullptr; }
  const XmlNode* path(std::initializer_list<const char*> p) const { const XmlNode* x = this; for (auto* n : p) { if (!x) return nullptr; x = x->kid(n); } return x; }
  std::vector<const XmlNode*> all(const char* n) const { std::vector<const XmlNode*> r; for (auto& c : kids) if (c->name == n) r.push_back(c.get()); return r; }
};

// elements, attributes, comments, declaration, CDATA-free text (ignored); entities &amp; &lt; &gt; &quot; &apos; in attribute values
struct XmlParser {
  const char* p; const char* e; std::string err;
  void ws() { while (p < e && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) p++; }
  bool starts(const char* s) const { size_t n = strlen(s); return (size_t)(e - p) >= n && memcmp(p, s, n) == 0; }
  void skip_misc() { for (;;) { ws(); if (starts("<?")) { const char* q = strstr(p, "?>"); p = q ? q + 2 : e; } else if (starts("<!--")) { const char* q = strstr(p, "-->"); p = q ? q + 3 : e; }
      else if (starts("<!")) { while (p < e && *p != '>') p++; if (p < e) p++; } else break; } }
  static std::string unescape(const std::string& s) { std::string o; for (size_t i = 0; i < s.size(); i++) { if (s[i] != '&') { o += s[i]; continue; }
      if (!s.compare(i, 5, "&amp;")) { o += '&'; i += 4; } else if (!s.compare(i, 4, "&lt;")) { o += '<'; i += 3; } else if (!s.compare(i, 4, "&gt;")) { o += '>'; i += 3; }
      else if (!s.compare(i, 6, "&quot;")) { o += '"'; i += 5; } else if (!s.compare(i, 6, "&apos;")) { o += '\''; i += 5; } else o += s[i]; } return o; }
  std::unique_ptr<XmlNode> element() {
    skip_misc();
    if (p >= e || *p != '<') { err = "element expected"; return nullptr; }
    p++;
    auto n = std::make_unique<XmlNode>();
    while (p < e && !strchr(" \t\r\n/>", *p)) n->name += *p++;
    for (;;) { ws(); if (p >= e) { err = "unterminated tag " + n->name; return nullptr; }
      if (*p == '/') { p++; if (p < e && *p == '>') { p++; return n; } err = "bad tag end"; return nullptr; }
      if (*p == '>') { p++; break; }
      std::string k; while (p < e && *p != '=' && !strchr(" \t\r\n", *p)) k += *p++;
      ws(); if (p >= e || *p != '=') { err = "attribute without value in " + n->name; return nullptr; } p++; ws();
      if (p >= e || (*p != '"' && *p != '\'')) { err = "unquoted attribute in " + n->name; return nullptr; }
      const char q = *p++; std::string v; while (p < e && *p != q) v += *p++; if (p < e) p++;
      n->attrs.push_back({k, unescape(v)}); }
    for (;;) { while (p < e && *p != '<') p++;                                             // text content is ignored
      if (p >= e) { err = "unterminated element " + n->name; return nullptr; }
      if (starts("</")) { while (p < e && *p != '>') p++; if (p < e) p++; return n; }
      if (starts("<!--") || starts("<?") || starts("<!")) { skip_misc(); continue; }
      auto c = element(); if (!c) return nullptr; n->kids.push_back(std::move(c)); }
  }
};

inline std::unique_ptr<XmlNode> xml_parse_file(const char* path, std::string& err) {
  FILE* f = fopen(path, "rb"); if (!f) { err = std::string("cannot open ") + path; return nullptr; }
  std::string buf; char tmp[65536]; size_t n; while ((n = fread(tmp, 1, sizeof tmp, f)) > 0) buf.append(tmp, n); fclose(f);
  XmlParser P{buf.data(), buf.data() + buf.size(), ""};
  auto r = P.element(); if (!r) err = "XML: " + P.err; return r;
}

inline double hms(const std::string& s) { int h = 0, m = 0; double sec = 0; if (sscanf(s.c_str(), "%d:%d:%lf", &h, &m, &sec) >= 2) return h * 3600.0 + m * 60.0 + sec; return strtod(s.c_str(), nullptr); }
inline std::vector<double> nums(const std::string& s) { std::vector<double> v; const char* p = s.c_str(); char* q; for (;;) { double x = strtod(p, &q); if (q == p) break; v.push_back(x); p = q; } return v; }

// ROOT_SYMUBRUIT -> flat tables.  Numbering: nodes in the order EXTREMITES, REPARTITEURS, CONVERGENTS, CARREFOURSAFEUX of the network
// section, links in TRONCONS order followed by the junction-internal links in their creation order.
inline bool flat_from_xml(const char* path, FlatNet& F, std::string& err) {
  auto root = xml_parse_file(path, err); if (!root) return false;
  if (root->name != "ROOT_SYMUBRUIT") { err = "root element is not ROOT_SYMUBRUIT"; return false; }
  const XmlNode* sim = root->path({"SIMULATIONS", "SIMULATION"}); const XmlNode* tra = root->path({"TRAFICS", "TRAFIC"}); const XmlNode* res = root->path({"RESEAUX", "RESEAU"});
  if (!sim || !tra || !res) { err = "SIMULATION / TRAFIC / RESEAU section missing"; return false; }
  if (sim->str("comportementflux", "iti") != "iti") { err = "only comportementflux=\"iti\" (predefined itineraries) is supported"; return false; }
  F = FlatNet();
  const double dt = sim->num("pasdetemps", 1.0);
  F.params.assign(SYMFLAT_P_COUNT, 0.0);
  F.params[SYMFLAT_P_DT] = dt; F.params[SYMFLAT_P_NSTEPS] = std::floor((hms(sim->str("fin", "01:00:00")) - hms(sim->str("debut", "00:00:00"))) / dt + 0.5);
  F.params[SYMFLAT_P_SEED] = sim->num("seed", 0); F.params[SYMFLAT_P_ACCBORNEE] = tra->flag("accbornee", false); F.params[SYMFLAT_P_RELAX] = tra->num("coeffrelax", 1.0);
  F.params[SYMFLAT_P_CHGTVOIE] = tra->flag("chgtvoie", false); F.params[SYMFLAT_P_DSTCHGT] = tra->num("chgtvoie_dstfin", 50.0);
  F.params[SYMFLAT_P_TRAVERSEES] = tra->flag("traversees", false); F.params[SYMFLAT_P_PHI_FORCE] = tra->num("chgtvoie_dstfin_force_phi", 1.0); F.params[SYMFLAT_P_DST_FORCE] = tra->num("chgtvoie_dstfin_force", 0.0);
  if (F.params[SYMFLAT_P_TRAVERSEES] != 0) { err = "traversees=\"true\" (conflict points inside junctions) is not supported"; return false; }
  // vehicle types (reseau.cpp:5640-5900)
  std::map<std::string, int> cls_id;
  if (auto* tv = tra->kid("TYPES_DE_VEHICULE")) for (auto* t : tv->all("TYPE_DE_VEHICULE")) { cls_id[t->str("id")] = (int)F.cls_names.size(); F.cls_names.push_back(t->str("id"));
    std::vector<double> r(SYMFLAT_CLS_STRIDE, 0.0); r[0] = t->num("w", -5.8823); r[1] = t->num("kx", 0.17); r[2] = t->num("vx", 25.0); r[3] = t->num("espacement_arret", 0.0);
    r[4] = t->num("deceleration", 0.0); r[5] = t->num("chgtvoie_tau", 4.0); r[6] = t->num("pi_rabattement", 0.0);
    const std::string loi = t->str("loi", "newell"); r[7] = loi == "gipps" ? SYMFLAT_LAW_GIPPS : loi == "idm" ? SYMFLAT_LAW_IDM : SYMFLAT_LAW_NEWELL;
    int np = 0; if (auto* ap = t->kid("ACCELERATION_PLAGES")) for (auto* a : ap->all("ACCELERATION_PLAGE")) if (np < 4) { r[9 + 2 * np] = a->num("ax", 1.5); r[10 + 2 * np] = a->num("vit_sup", 1e9); np++; }
    r[8] = np; F.cls.insert(F.cls.end(), r.begin(), r.end()); }
  const int T = (int)F.cls_names.size(); if (!T) { err = "no TYPE_DE_VEHICULE"; return false; }
  // nodes
  std::map<std::string, int> node_id; std::vector<const XmlNode*> node_xml; std::vector<double> node_xy;
  auto add_node = [&](const XmlNode* x, int type) { node_id[x->str("id")] = (int)F.node_names.size(); F.node_names.push_back(x->str("id")); F.node_i.insert(F.node_i.end(), {type, -1, 0, 0}); node_xml.push_back(x); node_xy.push_back(0); node_xy.push_back(0); };
  const XmlNode* cnx = res->kid("CONNEXIONS"); if (!cnx) { err = "RESEAU/CONNEXIONS missing"; return false; }
  if (const XmlNode* gi = cnx->kid("GIRATOIRES")) if (!gi->kids.empty()) { err = "GIRATOIRES (roundabouts) are not supported"; return false; }
  if (auto* g = cnx->kid("EXTREMITES")) for (auto* x : g->all("EXTREMITE")) add_node(x, 'E');            // 'E' or 'S' once the links are known
  if (auto* g = cnx->kid("REPARTITEURS")) for (auto* x : g->all("REPARTITEUR")) add_node(x, 'R');
  if (auto* g = cnx->kid("CONVERGENTS")) for (auto* x : g->all("CONVERGENT")) add_node(x, 'C');
  const int first_caf = (int)F.node_names.size();
  if (auto* g = cnx->kid("CARREFOURSAFEUX")) for (auto* x : g->all("CARREFOURAFEUX")) add_node(x, 'C');
  auto is_caf = [&](int n) { return n >= first_caf; };
  auto is_cvg = [&](int n) { return n < first_caf && F.node_i[4 * n] == 'C'; };
  // links (TuyauMicro, tuyau.cpp:1503-1615)
  std::map<std::string, int> link_id; std::vector<double> lane_w;
  const XmlNode* trs = res->kid("TRONCONS"); if (!trs) { err = "RESEAU/TRONCONS missing"; return false; }
  for (auto* t : trs->all("TRONCON")) { const int k = (int)F.link_names.size(); link_id[t->str("id")] = k; F.link_names.push_back(t->str("id"));
    auto a = node_id.find(t->str("id_eltamont")), b = node_id.find(t->str("id_eltaval"));
    if (a == node_id.end() || b == node_id.end()) { err = "TRONCON " + t->str("id") + ": unknown id_eltamont / id_eltaval"; return false; }
    const auto p0 = nums(t->str("extremite_amont", "0 0")), p1 = nums(t->str("extremite_aval", "0 0"));
    if (p0.size() < 2 || p1.size() < 2) { err = "TRONCON " + t->str("id") + ": bad coordinates"; return false; }
    const int nl = (int)t->num("nb_voie", 1); const double dx = p1[0] - p0[0], dy = p1[1] - p0[1];
    const double len = t->attr("longueur") ? t->num("longueur", 0) : std::sqrt(dx * dx + dy * dy);
    const int first = F.n_lanes();
    F.link_i.insert(F.link_i.end(), {first, nl, a->second, b->second, 0, -1, 0, 0}); F.link_f.push_back(len); F.link_f.push_back(t->num("vit_reg", 13.89));
    F.link_geom.insert(F.link_geom.end(), {p0[0], p0[1], p1[0], p1[1]}); lane_w.push_back(t->num("largeur_voie", 3.0));
    for (int n = 0; n < nl; n++) { F.lane_i.insert(F.lane_i.end(), {k, n, -1, 0}); F.lane_f.push_back(len); }
    if (F.node_i[4 * a->second] == 'E') { node_xy[2 * a->second] = p0[0]; node_xy[2 * a->second + 1] = p0[1]; }
    if (F.node_i[4 * b->second] == 'E') { F.node_i[4 * b->second] = 'S'; node_xy[2 * b->second] = p1[0]; node_xy[2 * b->second + 1] = p1[1]; }
    else { node_xy[2 * b->second] = p1[0]; node_xy[2 * b->second + 1] = p1[1]; } }
  const int n_ext_links = F.n_links();
  for (int k = 0; k < n_ext_links; k++) { const int dn = F.link_i[8 * k + 3]; const int t = F.node_i[4 * dn]; F.link_i[8 * k + 4] = t == 'S' ? 'S' : is_caf(dn) ? 'D' : t == 'C' ? 'C' : 'R'; }
  // signal controllers (ControleurDeFeux.cpp:203-307): one plan, its sequences in order
  std::map<std::string, int> ctrl_id;
  if (auto* cs = tra->kid("CONTROLEURS_DE_FEUX")) for (auto* c : cs->all("CONTROLEUR_DE_FEUX")) { ctrl_id[c->str("id")] = F.n_ctrl();
    const XmlNode* plan = c->path({"PLANS_DE_FEUX", "PLAN_DE_FEUX"}); double cycle = 0; const int seq0 = (int)F.seq_i.size() / 2; int nseq = 0;
    if (plan) if (auto* ss = plan->kid("SEQUENCES")) for (auto* s : ss->all("SEQUENCE")) { const double d = s->num("duree_totale", 0); cycle += d; nseq++;
      const int g0 = (int)F.sig_i.size() / 2; int ng = 0;
      if (auto* sa = s->kid("SIGNAUX_ACTIFS")) for (auto* g : sa->all("SIGNAL_ACTIF")) { auto li = link_id.find(g->str("troncon_entree")), lo = link_id.find(g->str("troncon_sortie"));
        if (li == link_id.end() || lo == link_id.end()) { err = "SIGNAL_ACTIF: unknown link"; return false; }
        F.sig_i.push_back(li->second); F.sig_i.push_back(lo->second); F.sig_f.insert(F.sig_f.end(), {g->num("duree_vert", 0), g->num("duree_orange", 0), g->num("duree_retard_allumage", 0)}); ng++; }
      F.seq_i.push_back(g0); F.seq_i.push_back(ng); F.seq_f.push_back(d); }
    F.ctrl_i.push_back(seq0); F.ctrl_i.push_back(nseq); F.ctrl_f.push_back(cycle); }
  for (size_t n = 0; n < node_xml.size(); n++) if (auto* v = node_xml[n]->attr("controleur_de_feux")) { auto it = ctrl_id.find(*v); if (it == ctrl_id.end()) { err = "unknown controleur_de_feux " + *v; return false; } F.node_i[4 * n + 1] = it->second; }
  // allowed movements (Connection.h m_mapMvtAutorises): per connection, entry lane -> (exit link, exit lane, rate, stop-line offset)
  struct Mv { int a, ka, o, ko; double taux, ligne; };
  std::vector<std::vector<Mv>> node_mv(F.node_names.size());
  std::map<std::string, const XmlNode*> ci_xml;
  if (auto* ci = tra->kid("CONNEXIONS_INTERNES")) for (auto* c : ci->all("CONNEXION_INTERNE")) { ci_xml[c->str("id")] = c; auto nit = node_id.find(c->str("id")); if (nit == node_id.end()) continue;
    if (auto* ma = c->kid("MOUVEMENTS_AUTORISES")) for (auto* m : ma->all("MOUVEMENT_AUTORISE")) { auto la = link_id.find(m->str("id_troncon_amont")); if (la == link_id.end()) { err = "MOUVEMENT_AUTORISE: unknown upstream link"; return false; }
      const int ka = (int)m->num("num_voie_amont", 1) - 1;
      if (auto* ms = m->kid("MOUVEMENT_SORTIES")) for (auto* s : ms->all("MOUVEMENT_SORTIE")) { auto lo = link_id.find(s->str("id_troncon_aval")); if (lo == link_id.end()) { err = "MOUVEMENT_SORTIE: unknown downstream link"; return false; }
        node_mv[nit->second].push_back({la->second, ka, lo->second, (int)s->num("num_voie_aval", 1) - 1, s->num("taux_utilisation", 1.0), s->num("ligne_feu", 0.0)}); } } }
  std::vector<std::vector<std::array<double, 4>>> succ(F.n_lanes());      // per lane: key link, next lane, flags, rate
  std::vector<std::vector<std::array<double, 6>>> sl(F.n_lanes());        // per lane: ctrl, up link, up num, down link, down num, pos
  auto lane_of = [&](int link, int num) { return F.link_i[8 * link] + num; };
  auto lane_pt = [&](int link, int num, bool end, double& px, double& py) { const double* g = &F.link_geom[4 * link]; const double dx = g[2] - g[0], dy = g[3] - g[1]; const double L = std::sqrt(dx * dx + dy * dy);
    const double ux = L > 0 ? dx / L : 1, uy = L > 0 ? dy / L : 0; const double off = (num + 0.5) * lane_w[link]; px = (end ? g[2] : g[0]) + uy * off; py = (end ? g[3] : g[1]) - ux * off; };
  std::vector<std::vector<int>> lcam(F.n_links()), lcav(F.n_links());
  struct Cv { std::string name; int node, down; double tagreg, gamma, mu, pos; std::vector<double> tf; std::vector<int> am; std::vector<std::vector<std::pair<int, int>>> vam; };
  std::vector<Cv> cvs;
  const double def_tagreg = tra->num("PeriodeAgregationCapteurs", 30.0), def_gamma = tra->num("Gamma", 1.0), def_mu = tra->num("Mu", 1.0), def_ti = tra->num("ti", 0.0);
  auto tf_of = [&](const XmlNode* ci) { std::vector<double> tf(T, def_ti); if (ci) if (auto* l = ci->kid("LISTE_TEMPS_INSERTION")) for (auto& x : l->kids) { auto it = cls_id.find(x->str("id_typevehicule")); if (it != cls_id.end()) tf[it->second] = x->num("ti", def_ti); } return tf; };
  auto add_internal = [&](const std::string& name, int node, double len, int type_aval, int prev_lane, double x0, double y0, double x1, double y1, double vreg) {
    const int k = F.n_links(); const int first = F.n_lanes(); F.link_names.push_back(name);
    F.link_i.insert(F.link_i.end(), {first, 1, node, node, type_aval, node, 0, 0}); F.link_f.push_back(len); F.link_f.push_back(vreg);
    F.link_geom.insert(F.link_geom.end(), {x0, y0, x1, y1}); lane_w.push_back(3.0);
    F.lane_i.insert(F.lane_i.end(), {k, 0, prev_lane, 0}); F.lane_f.push_back(len); succ.emplace_back(); sl.emplace_back(); lcam.emplace_back(); lcav.emplace_back();
    return k; };
  for (size_t n = 0; n < node_mv.size(); n++) { auto& mv = node_mv[n]; if (mv.empty()) continue;
    const int ctrl = F.node_i[4 * n + 1];
    if (!is_caf((int)n)) {                                                                 // distributor or plain merge: movements are lane -> lane
      for (auto& m : mv) { const int la = lane_of(m.a, m.ka); succ[la].push_back({(double)m.o, (double)lane_of(m.o, m.ko), 1.0, m.taux});
        if (ctrl >= 0) sl[la].push_back({(double)ctrl, (double)m.a, (double)m.ka, (double)m.o, (double)m.ko, F.lane_f[la] + m.ligne});
        if (std::find(lcav[m.a].begin(), lcav[m.a].end(), m.o) == lcav[m.a].end()) lcav[m.a].push_back(m.o);
        if (!is_cvg((int)n) && std::find(lcam[m.o].begin(), lcam[m.o].end(), m.a) == lcam[m.o].end()) lcam[m.o].push_back(m.a); }
      continue; }
    // signalised junction: internal links in the creation order of CarrefourAFeuxEx::Init (entry-lane label, exit-link label, exit lane)
    std::vector<int> idx(mv.size()); for (size_t i = 0; i < mv.size(); i++) idx[i] = (int)i;
    auto key = [&](const Mv& m) { return F.link_names[m.a] + "V" + std::to_string(m.ka + 1); };
    std::stable_sort(idx.begin(), idx.end(), [&](int x, int y) { const Mv &A = mv[x], &B = mv[y]; const std::string ka = key(A), kb = key(B); if (ka != kb) return ka < kb;
      if (F.link_names[A.o] != F.link_names[B.o]) return F.link_names[A.o] < F.link_names[B.o]; return A.ko < B.ko; });
    std::map<int, int> feeds_count; for (auto& m : mv) feeds_count[m.o]++;
    const double vmax = node_xml[n]->num("vit_max", 1e9);
    std::vector<int> exits; std::vector<std::vector<std::pair<int, int>>> feeds;
    for (int i : idx) { const Mv& m = mv[i]; double ax, ay, ox, oy; lane_pt(m.a, m.ka, true, ax, ay); lane_pt(m.o, m.ko, false, ox, oy);
      const double dx = ox - ax, dy = oy - ay; const double len = std::sqrt(dx * dx + dy * dy); const int ent = lane_of(m.a, m.ka);
      const int il = add_internal("I" + F.link_names[m.a] + "_" + F.link_names[m.o] + "_" + std::to_string(m.ka + 1), (int)n, len, feeds_count[m.o] > 1 ? 'C' : 'R', ent, ax, ay, ox, oy,
                                  vmax);                                                    // the junction's vit_max (CarrefourAFeuxEx.cpp:480, :697)
      const int ilane = F.link_i[8 * il];
      succ[ent].push_back({(double)m.o, (double)ilane, 1.0, m.taux}); succ[ilane].push_back({(double)m.o, (double)lane_of(m.o, m.ko), 1.0, 1.0});
      if (ctrl >= 0) { if (m.ligne <= 0) sl[ent].push_back({(double)ctrl, (double)m.a, (double)m.ka, (double)m.o, (double)m.ko, F.lane_f[ent] + m.ligne});
                       else sl[ilane].push_back({(double)ctrl, (double)m.a, (double)m.ka, (double)m.o, (double)m.ko, m.ligne}); }       // CarrefourAFeuxEx::FinInit :1205-1275
      size_t xi = 0; while (xi < exits.size() && exits[xi] != m.o) xi++;
      if (xi == exits.size()) { exits.push_back(m.o); feeds.emplace_back(); }
      feeds[xi].push_back({il, m.ko});
      F.cafmvt_i.insert(F.cafmvt_i.end(), {(int)n, m.a, m.o, (int)F.cafmvt_links.size(), 1}); F.cafmvt_links.push_back(il);
      lcam[il] = {m.a}; lcav[m.a].push_back(il); lcav[il] = {m.o}; }
    const XmlNode* ci = ci_xml.count(F.node_names[n]) ? ci_xml[F.node_names[n]] : nullptr;
    for (size_t xi = 0; xi < exits.size(); xi++) { const int o = exits[xi]; for (auto& fd : feeds[xi]) lcam[o].push_back(fd.first);
      if (feeds[xi].size() < 2) continue;                                                  // one upstream link: a distributor (CarrefourAFeuxEx.cpp:860-905)
      Cv cv; cv.name = F.node_names[n] + "_C0_" + F.link_names[o]; cv.node = (int)n; cv.down = o; cv.pos = 5.0;
      cv.tagreg = ci ? ci->num("PeriodeAgregationCapteurs", def_tagreg) : def_tagreg; cv.gamma = ci ? ci->num("Gamma", def_gamma) : def_gamma; cv.mu = ci ? ci->num("Mu", def_mu) : def_mu; cv.tf = tf_of(ci);
      for (auto& fd : feeds[xi]) cv.am.push_back(fd.first);
      for (int kk = 0; kk < F.link_i[8 * o + 1]; kk++) { double vx0, vy0, vx1, vy1; lane_pt(o, kk, false, vx0, vy0); lane_pt(o, kk, true, vx1, vy1);
        std::vector<std::pair<double, int>> ang; std::vector<int> lanes_kk;
        for (auto& fd : feeds[xi]) if (fd.second == kk) { const int ln = F.link_i[8 * fd.first]; lanes_kk.push_back(ln); const double* g = &F.link_geom[4 * fd.first];
          const double a1 = vx1 - vx0, a2 = vy1 - vy0, b1 = g[0] - vx0, b2 = g[1] - vy0; double t = ((a1 * a1 + a2 * a2) != 0 && (b1 * b1 + b2 * b2) != 0) ? std::atan2(a1 * b2 - a2 * b1, a1 * b1 + a2 * b2) : 0.0;
          if (t < 0) t += 2 * 3.14159265358979323846; ang.push_back({t, ln}); }                // AngleOfView tools.cpp:1323-1365
        std::sort(ang.begin(), ang.end());
        std::vector<std::pair<int, int>> v; for (int ln : lanes_kk) { int lev = 0; for (size_t q = 0; q < ang.size(); q++) if (ang[q].second == ln) lev = (int)ang.size() - (int)q; v.push_back({ln, lev}); }   // :1129-1166
        cv.vam.push_back(v); }
      cvs.push_back(std::move(cv)); }
  }
  // plain merges (reseau.cpp:6432-6473, :8300-8420): every upstream lane is an upstream lane of every point
  if (auto* g = cnx->kid("CONVERGENTS")) for (auto* x : g->all("CONVERGENT")) { const int n = node_id[x->str("id")]; auto lo = link_id.find(x->str("id_TAv")); if (lo == link_id.end()) { err = "CONVERGENT " + x->str("id") + ": unknown id_TAv"; return false; }
    const XmlNode* ci = ci_xml.count(x->str("id")) ? ci_xml[x->str("id")] : nullptr;
    Cv cv; cv.name = x->str("id"); cv.node = -1; cv.down = lo->second; cv.pos = ci ? ci->num("pos_cpt_Av", tra->num("pos_cpt_Av", 5.0)) : tra->num("pos_cpt_Av", 5.0);
    cv.tagreg = ci ? ci->num("PeriodeAgregationCapteurs", def_tagreg) : def_tagreg; cv.gamma = ci ? ci->num("Gamma", def_gamma) : def_gamma; cv.mu = ci ? ci->num("Mu", def_mu) : def_mu; cv.tf = tf_of(ci);
    std::vector<std::pair<int, int>> up;
    if (auto* ta = x->kid("TRONCONS_AMONT")) for (auto* a : ta->all("TRONCON_AMONT")) { auto la = link_id.find(a->str("id")); if (la == link_id.end()) { err = "TRONCON_AMONT: unknown link"; return false; }
      cv.am.push_back(la->second); for (int k = 0; k < F.link_i[8 * la->second + 1]; k++) up.push_back({lane_of(la->second, k), (int)a->num("priorite", 1)}); }
    for (int l = 0; l < F.link_i[8 * lo->second + 1]; l++) cv.vam.push_back(up);
    lcam[lo->second] = cv.am; (void)n;
    cvs.push_back(std::move(cv)); }
  // CSR tables
  const int L = F.n_lanes(), K = F.n_links();
  F.succ_off.assign(L + 1, 0); F.sl_off.assign(L + 1, 0);
  for (int l = 0; l < L; l++) { for (auto& s : succ[l]) { F.succ_i.insert(F.succ_i.end(), {(int)s[0], (int)s[1], (int)s[2]}); F.succ_f.push_back(s[3]); }
    F.succ_off[l + 1] = (int)F.succ_f.size();
    for (auto& s : sl[l]) { F.sl_i.insert(F.sl_i.end(), {(int)s[0], (int)s[1], (int)s[2], (int)s[3], (int)s[4], 0}); F.sl_f.push_back(s[5]); }
    F.sl_off[l + 1] = (int)F.sl_f.size(); }
  std::sort(cvs.begin(), cvs.end(), [](const Cv& a, const Cv& b) { return a.name < b.name; });   // Reseau::Liste_convergents: std::map by id
  for (size_t c = 0; c < cvs.size(); c++) { const Cv& cv = cvs[c];
    F.cvg_i.insert(F.cvg_i.end(), {cv.node, cv.down, F.n_pt(), (int)cv.vam.size(), (int)F.cvgam_i.size(), (int)cv.am.size(), 0, 0}); F.cvg_f.insert(F.cvg_f.end(), {cv.tagreg, cv.gamma, cv.mu, cv.pos});
    F.cvg_tf.insert(F.cvg_tf.end(), cv.tf.begin(), cv.tf.end()); F.cvgam_i.insert(F.cvgam_i.end(), cv.am.begin(), cv.am.end());
    for (size_t l = 0; l < cv.vam.size(); l++) { F.pt_i.insert(F.pt_i.end(), {(int)c, lane_of(cv.down, (int)l), (int)F.ptam_i.size() / 2, (int)cv.vam[l].size()}); for (auto& a : cv.vam[l]) { F.ptam_i.push_back(a.first); F.ptam_i.push_back(a.second); } } }
  F.lcam_off.assign(K + 1, 0); F.lcav_off.assign(K + 1, 0);
  for (int k = 0; k < K; k++) { F.lcam_links.insert(F.lcam_links.end(), lcam[k].begin(), lcam[k].end()); F.lcam_off[k + 1] = (int)F.lcam_links.size();
    F.lcav_links.insert(F.lcav_links.end(), lcav[k].begin(), lcav[k].end()); F.lcav_off[k + 1] = (int)F.lcav_links.size(); }
  // predefined routes
  std::map<std::string, int> route_id; F.route_off.assign(1, 0);
  if (auto* rs = res->kid("ROUTES")) for (auto* r : rs->all("ROUTE")) { route_id[r->str("id")] = (int)F.route_off.size() - 1;
    if (auto* ts = r->kid("TRONCONS_")) for (auto* t : ts->all("TRONCON_")) { auto it = link_id.find(t->str("id")); if (it == link_id.end()) { err = "ROUTE " + r->str("id") + ": unknown link"; return false; } F.route_links.push_back(it->second); }
    F.route_off.push_back((int)F.route_links.size()); }
  // exits (sortie.cpp:60-130) first: the origins refer to them; then origins in document order (Reseau::Liste_origines)
  std::map<std::string, int> exit_id;
  const XmlNode* ext = tra->kid("EXTREMITES");
  std::map<std::string, const XmlNode*> ext_xml; if (ext) for (auto* x : ext->all("EXTREMITE")) ext_xml[x->str("id")] = x;
  for (size_t n = 0; n < F.node_names.size(); n++) if (F.node_i[4 * n] == 'S') { int link = -1; for (int k = 0; k < n_ext_links; k++) if (F.link_i[8 * k + 3] == (int)n) link = k;
    double cap = -1.0; auto it = ext_xml.find(F.node_names[n]); if (it != ext_xml.end()) if (auto* c = it->second->path({"CAPACITES", "CAPACITE"})) cap = c->num("valeur", -1.0);
    exit_id[F.node_names[n]] = F.n_exits(); F.exit_i.push_back((int)n); F.exit_i.push_back(link); F.exit_f.push_back(cap); }
  if (ext) for (auto* x : ext->all("EXTREMITE")) { auto nit = node_id.find(x->str("id")); if (nit == node_id.end() || F.node_i[4 * nit->second] != 'E') continue;
    const XmlNode* fl = x->path({"FLUX_GLOBAL", "FLUX"}); if (!fl) continue;
    int link = -1; for (int k = 0; k < n_ext_links; k++) if (F.link_i[8 * k + 2] == nit->second) link = k;
    if (link < 0) { err = "origin " + x->str("id") + " without a link"; return false; }
    const int dem0 = (int)F.dem_f.size() / 2; int nd = 0; if (auto* ds = fl->kid("DEMANDES")) for (auto* d : ds->all("DEMANDE")) { F.dem_f.push_back(d->num("duree", 1e12)); F.dem_f.push_back(d->num("niveau", 0)); nd++; }
    const int dest0 = (int)F.dest_f.size(); int ndest = 0;
    if (auto* rd = fl->path({"REP_DESTINATIONS", "REP_DESTINATION"})) for (auto* d : rd->all("DESTINATION")) { auto xit = exit_id.find(d->str("sortie")); if (xit == exit_id.end()) { err = "DESTINATION: unknown exit " + d->str("sortie"); return false; }
      const int r0 = (int)F.droute_i.size(); int nr = 0; for (auto* r : d->all("ROUTE")) { auto rit = route_id.find(r->str("id")); if (rit == route_id.end()) { err = "DESTINATION: unknown route " + r->str("id"); return false; }
        F.droute_i.push_back(rit->second); F.droute_f.push_back(r->num("coeffAffectation", 1.0)); nr++; }
      F.dest_i.insert(F.dest_i.end(), {xit->second, r0, nr}); F.dest_f.push_back(d->num("coeffOD", 1.0)); ndest++; }
    const int nl = F.link_i[8 * link + 1]; const int lc0 = (int)F.lanecoef_f.size(); std::vector<double> lc; if (auto* rv = fl->path({"REP_VOIES", "REP_VOIE"})) lc = nums(rv->str("coeffs"));
    for (int k = 0; k < nl; k++) F.lanecoef_f.push_back(k < (int)lc.size() ? lc[k] : 1.0 / nl);
    const int tc0 = (int)F.typecoef_f.size(); std::vector<double> tc; if (auto* rt = fl->path({"REP_TYPEVEHICULES", "REP_TYPEVEHICULE"})) tc = nums(rt->str("coeffs"));
    for (int k = 0; k < T; k++) F.typecoef_f.push_back(k < (int)tc.size() ? tc[k] : (k == 0 ? 1.0 : 0.0));
    const bool expo = x->str("typeCreationVehicule", "demande") == "distributionExponentielle";
    F.entry_i.insert(F.entry_i.end(), {nit->second, link, dem0, nd, dest0, ndest, lc0, tc0 | ((int)expo << 30)}); }
  // initial vehicles (reseau.cpp:11680-11773)
  if (auto* tj = res->path({"INIT", "TRAJS"})) for (auto* t : tj->all("TRAJ")) { auto lk = link_id.find(t->str("tron")); auto rt = route_id.find(t->str("route")); auto ct = cls_id.find(t->str("type_veh"));
    if (lk == link_id.end() || rt == route_id.end() || ct == cls_id.end()) { err = "INIT/TRAJ " + t->str("id") + ": unknown link, route or vehicle type"; return false; }
    F.iv_i.insert(F.iv_i.end(), {ct->second, lane_of(lk->second, (int)t->num("voie", 1) - 1), rt->second, 0}); F.iv_f.insert(F.iv_f.end(), {t->num("dst", 0), t->num("vit", 0), t->num("acc", 0)});
    if (t->attr("id")) F.iv_id.push_back((int)t->num("id", 0)); }
  if (!F.iv_id.empty() && F.iv_id.size() != F.iv_i.size() / 4) F.iv_id.clear();
  F.node_xy = node_xy;
  return F.validate();
}

}  // namespace symb200
