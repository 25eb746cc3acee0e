// symexp.cpp — SymuVia's C exports (include/symuvia_exports.h) on top of the symgpu C-ABI.
// Mirrors symuvia/src/Symubruit.cpp:4239-4598 (extern "C" wrappers) and :142-242 (_SymRunNextStep):
// one global current network (theApp map with DEFAULT_NETWORK_ID, Symubruit.cpp:67,101), lazy start on the
// first step, bool/int error returns only, <INST> fragment copied into the caller's buffer with strcpy.
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/symuvia_exports.h"
#include "engine_internal.h"
#include "xmlnet.h"

namespace {
symgpu_ctx* g_net = nullptr;          // theApp.GetNetwork(DEFAULT_NETWORK_ID)
std::string g_last_inst;

bool ends_with(const std::string& s, const char* suf) { size_t n = strlen(suf); return s.size() >= n && s.compare(s.size() - n, n, suf) == 0; }

// SystemUtil::ToString(2, x): fixed notation, 2 decimals (XMLDocTrafic.cpp:2232-2257)
void fmt2(std::string& out, double x) { char b[48]; snprintf(b, sizeof b, "%.2f", x); out += b; }

// <INST> fragment: XMLDocTrafic::AddInstant (XMLDocTrafic.cpp:345-395) + AddTrajectoire (:2210-2330)
void build_inst(symgpu_ctx* c, std::string& out, int n, const std::vector<int>& id, const std::vector<int>& lane, const std::vector<double>& pos,
                const std::vector<double>& vit, const std::vector<double>& acc) {
  const symb200::FlatNet& f = symgpu_internal_net(c);
  out.clear();
  out.reserve(160 + (size_t)n * 120);
  char b[96];
  snprintf(b, sizeof b, "<INST val=\"%.2f\" nbVeh=\"%d\"><CREATIONS/><SORTIES>", symgpu_time(c), n); out += b;
  int nex = symgpu_get_exited(c, 0, nullptr, nullptr);
  if (nex > 0) { std::vector<int> xi(nex); std::vector<double> xt(nex); symgpu_get_exited(c, nex, xi.data(), xt.data());
    for (int k = 0; k < nex; k++) { snprintf(b, sizeof b, "<SORTIE id=\"%d\" instant=\"%.2f\"/>", xi[k], xt[k]); out += b; } }
  out += "</SORTIES><TRAJS>";
  for (int k = 0; k < n; k++) {
    if (pos[k] < 0) continue;                       // created but not yet on the network (UNDEF_POSITION)
    int ln = lane[k], link = f.lane_i[4 * ln], num = f.lane_i[4 * ln + 1];
    double len = f.lane_f[ln], fr = len > 0 ? pos[k] / len : 0;
    double x = 0, y = 0;
    if (!f.link_geom.empty()) { const double* g = &f.link_geom[4 * link]; x = g[0] + fr * (g[2] - g[0]); y = g[1] + fr * (g[3] - g[1]); }
    out += "<TRAJ id=\""; out += std::to_string(id[k]); out += "\" tron=\"";
    out += link < (int)f.link_names.size() ? f.link_names[link] : std::to_string(link);
    out += "\" voie=\""; out += std::to_string(num + 1);
    out += "\" dst=\""; fmt2(out, pos[k]); out += "\" abs=\""; fmt2(out, x); out += "\" ord=\""; fmt2(out, y); out += "\" z=\"0.00\" vit=\"";
    fmt2(out, vit[k]); out += "\" acc=\""; fmt2(out, acc[k]); out += "\" type=\"";
    int cl = symgpu_internal_class_of(c, id[k]); out += cl >= 0 && cl < (int)f.cls_names.size() ? f.cls_names[cl] : std::to_string(cl);
    out += "\"/>";
  }
  out += "</TRAJS><STREAMS/><LINKS/><SGTS/><FEUX/><ENTREES/><REGULATIONS/></INST>";
}

// the rows come packed, in m_LstVehicles order, through the trajectory path of the C-ABI (one copy per field on the side stream)
std::vector<int> g_id, g_lane; std::vector<double> g_pos, g_vit, g_acc;
bool run_next_step(char* out_xml, bool* end, bool release_at_end) {
  if (!g_net) return false;                                         // Symubruit.cpp:146-150
  int r;
  if (out_xml) { const size_t cap = (size_t)symgpu_num_vehicles(g_net) + 65536;
    if (g_id.size() < cap) { g_id.resize(cap); g_lane.resize(cap); g_pos.resize(cap); g_vit.resize(cap); g_acc.resize(cap); }
    r = symgpu_step_traj(g_net, (int)g_id.size(), g_id.data(), g_lane.data(), g_pos.data(), g_vit.data(), g_acc.data()); }
  else r = symgpu_step(g_net);
  if (r < 0) return false;
  const symb200::FlatNet& f = symgpu_internal_net(g_net);
  bool fin = symgpu_time(g_net) >= f.params[SYMFLAT_P_NSTEPS] * f.params[SYMFLAT_P_DT] - 1e-9;   // reseau.cpp:2427
  if (out_xml) { build_inst(g_net, g_last_inst, r, g_id, g_lane, g_pos, g_vit, g_acc); strcpy(out_xml, g_last_inst.c_str()); }   // Symubruit.cpp:1013
  if (end) *end = fin;
  if (fin && release_at_end) { symgpu_destroy(g_net); g_net = nullptr; }                        // Symubruit.cpp:1010-1011: only the char* overload releases the networks
  return true;
}
}  // namespace

extern "C" {

int SymLoadNetworkEx(char* sFile) {
  if (!sFile) return 0;
  if (g_net) { symgpu_destroy(g_net); g_net = nullptr; }
  // a ROOT_SYMUBRUIT document (the subset of csrc/xmlnet.h) or an already flattened *.symflat file: symgpu_create reads both
  symgpu_ctx* c = nullptr;
  if (symgpu_create(sFile, 0, &c) != SYMGPU_OK) { fprintf(stderr, "SymLoadNetworkEx: %s\n", symgpu_last_error()); return 0; }
  g_net = c;
  return 1;
}

void SymUnloadCurrentNetworkEx(void) { if (g_net) { symgpu_destroy(g_net); g_net = nullptr; } }

bool SymRunNextStepEx(char* sOutputXmlFlow, bool bTrace, bool* bNEnd) { (void)bTrace; return run_next_step(sOutputXmlFlow, bNEnd, true); }

bool SymRunNextStepLiteEx(bool bTrace, bool* bNEnd) { (void)bTrace; return run_next_step(nullptr, bNEnd, false); }   // Symubruit.cpp:1051-1060: the network stays loaded

int SymRunEx(char* sFile) {
  if (!SymLoadNetworkEx(sFile)) return 1;
  bool end = false;
  while (!end) if (!run_next_step(nullptr, &end, false)) { SymUnloadCurrentNetworkEx(); return 1; }
  SymUnloadCurrentNetworkEx();
  return 0;
}

int SymCreateVehicleEx(char* sType, char* sOrigin, char* sDestination, int nLane, double dbt) {
  if (!g_net) return -1;                                            // Symubruit.cpp:1725-1729
  const symb200::FlatNet& f = symgpu_internal_net(g_net);
  int cls = -1, entry = -1, dest = -1;
  for (size_t i = 0; i < f.cls_names.size(); i++) if (sType && f.cls_names[i] == sType) cls = (int)i;
  if (cls < 0) return -2;                                           // reseau.cpp:13380-13382
  for (int e = 0; e < f.n_entries(); e++) { int node = f.entry_i[8 * e]; if (sOrigin && node < (int)f.node_names.size() && f.node_names[node] == sOrigin) entry = e; }
  if (entry < 0) return -3;                                         // reseau.cpp:13440-13441
  const int* er = &f.entry_i[8 * entry];
  int nl = f.link_i[8 * er[1] + 1];
  if (nLane - 1 != -1 && nl <= nLane - 1) return -5;                // reseau.cpp:13444-13449 (lane numbers start at 1)
  for (int d = er[4]; d < er[4] + er[5]; d++) { int node = f.exit_i[2 * f.dest_i[3 * d]]; if (sDestination && node < (int)f.node_names.size() && f.node_names[node] == sDestination) dest = d - er[4]; }
  if (dest < 0) return -4;
  return symgpu_internal_create_vehicle(g_net, cls, entry, dest, nLane - 1, dbt);
}

void SymRunDeleteStr(char* str) { delete[] str; }                   // Symubruit.cpp:2303

// Host-only helper (no CUDA): flattens a ROOT_SYMUBRUIT document into a SYMFLAT1 file, so that the flattening the XML front end does
// (internal junction links, stop lines, merge points and their priorities) can be inspected and fed to other tools.  0 = ok.
int symflat_from_xml(const char* xml_path, const char* out_path) {
  if (!xml_path || !out_path) return 1;
  symb200::FlatNet f; std::string err;
  if (!symb200::flat_from_xml(xml_path, f, err)) { fprintf(stderr, "symflat_from_xml: %s\n", err.c_str()); return 1; }
  return f.save(out_path) ? 0 : 1;
}

}  // extern "C"
