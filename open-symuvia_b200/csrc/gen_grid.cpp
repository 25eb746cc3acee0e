// gen_grid.cpp — fast generator of the pre-seeded signalised Manhattan grids used by bench.py (BASELINE configs C4 / C5).
// Same network model as open-symuvia_b200/synth.py:grid() (CAF junctions flattened into one internal lane per movement,
// CarrefourAFeuxEx.cpp:660-732; stop lines at the entry lane ends, :1205-1275; two-sequence fixed-time plan), written in
// C++ because an (nx = 800) x 100 grid has 2.5 M lanes.  Host-only, no CUDA.  Every rank of a partitioned run generates the
// full network but only the vehicles of its own strip; vehicle ids are global (prefix sum of per-lane counts) so that the
// union over ranks is one population listed in id order.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace {
struct Rng { uint64_t s; explicit Rng(uint64_t seed) : s(seed * 0x9E3779B97F4A7C15ull + 0x1234567ull) {}
  uint64_t next() { uint64_t z = (s += 0x9E3779B97F4A7C15ull); z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }
  double uni() { return (next() >> 11) * (1.0 / 9007199254740992.0); } };

struct Writer { FILE* f; int n = 0; long npos;
  explicit Writer(const char* p) { f = fopen(p, "wb"); if (f) { fwrite("SYMFLAT1", 1, 8, f); npos = ftell(f); uint32_t z = 0; fwrite(&z, 4, 1, f); } }
  void sec(const char* name, uint32_t dt, const void* data, uint64_t count) { char nm[24] = {0}; strncpy(nm, name, 23); fwrite(nm, 1, 24, f); fwrite(&dt, 4, 1, f); fwrite(&count, 8, 1, f);
    size_t isz = dt == 0 ? 4 : dt == 1 ? 8 : 1, bytes = count * isz; if (bytes) fwrite(data, 1, bytes, f); size_t pad = (8 - bytes % 8) % 8; char z[8] = {0}; if (pad) fwrite(z, 1, pad, f); n++; }
  void I(const char* name, const std::vector<int>& v) { sec(name, 0, v.data(), v.size()); }
  void D(const char* name, const std::vector<double>& v) { sec(name, 1, v.data(), v.size()); }
  bool close() { if (!f) return false; fseek(f, npos, SEEK_SET); uint32_t k = (uint32_t)n; fwrite(&k, 4, 1, f); fclose(f); f = nullptr; return true; } };
const int DX[4] = {1, 0, -1, 0}, DY[4] = {0, 1, 0, -1};
}  // namespace

extern "C" int symgen_grid(const char* path, int nx, int ny, double per_lane, int seed, int route_len, int n_steps, int rank, int world,
                           long long* n_vehicles_total) {
  const double block = 200.0, setback = 12.0, lane_w = 3.0, vit_reg = 13.89, green = 30.0, amber = 3.0;
  const int NL = 2;
  const double w = -5.8823, kx = 0.17, vx = 25.0;
  const double L = block - 2 * setback;
  auto J = [&](int i, int j) { return i * ny + j; };
  std::vector<int> node_i; std::vector<double> node_xy;
  for (int i = 0; i < nx; i++) for (int j = 0; j < ny; j++) { node_i.insert(node_i.end(), {(int)'C', J(i, j), 0, 0}); node_xy.push_back(i * block); node_xy.push_back(j * block); }
  std::vector<int> link_i, lane_i; std::vector<double> link_f, link_geom, lane_f;
  std::vector<int> link_dir, link_kind;                      // kind: 0 'L' junction->junction, 1 exit stub, 2 entry stub, 3 internal
  std::vector<int> out_link((size_t)nx * ny * 4, -1), in_link((size_t)nx * ny * 4, -1);
  std::vector<int> exit_i; std::vector<double> exit_f;
  auto add_link = [&](int up, int down, int nl, double len, char ta, int brick, int prev_lane, double x0, double y0, double x1, double y1, int dir, int kind) {
    int k = (int)link_i.size() / 8; int first = (int)lane_i.size() / 4;
    link_i.insert(link_i.end(), {first, nl, up, down, (int)ta, brick, 0, 0}); link_f.push_back(len); link_f.push_back(vit_reg);
    link_geom.insert(link_geom.end(), {x0, y0, x1, y1}); link_dir.push_back(dir); link_kind.push_back(kind);
    for (int n = 0; n < nl; n++) { lane_i.insert(lane_i.end(), {k, n, prev_lane, 0}); lane_f.push_back(len); }
    return k; };
  for (int i = 0; i < nx; i++) for (int j = 0; j < ny; j++) {
    double x0 = i * block, y0 = j * block;
    for (int d = 0; d < 4; d++) {
      int i2 = i + DX[d], j2 = j + DY[d]; double gx0 = x0 + DX[d] * setback, gy0 = y0 + DY[d] * setback;
      if (i2 >= 0 && i2 < nx && j2 >= 0 && j2 < ny) {
        int k = add_link(J(i, j), J(i2, j2), NL, L, 'D', -1, -1, gx0, gy0, i2 * block - DX[d] * setback, j2 * block - DY[d] * setback, d, 0);
        out_link[(size_t)J(i, j) * 4 + d] = k; in_link[(size_t)J(i2, j2) * 4 + d] = k;
      } else {
        int s = (int)node_i.size() / 4; node_i.insert(node_i.end(), {(int)'S', -1, 0, 0}); node_xy.push_back(x0 + DX[d] * block); node_xy.push_back(y0 + DY[d] * block);
        int k = add_link(J(i, j), s, NL, L, 'S', -1, -1, gx0, gy0, x0 + DX[d] * (block - setback), y0 + DY[d] * (block - setback), d, 1);
        out_link[(size_t)J(i, j) * 4 + d] = k; exit_i.push_back(s); exit_i.push_back(k); exit_f.push_back(-1.0);
        int e = (int)node_i.size() / 4; node_i.insert(node_i.end(), {(int)'E', -1, 0, 0}); node_xy.push_back(x0 + DX[d] * block); node_xy.push_back(y0 + DY[d] * block);
        int din = (d + 2) % 4;
        int k2 = add_link(e, J(i, j), NL, L, 'D', -1, -1, x0 + DX[d] * (block - setback), y0 + DY[d] * (block - setback), gx0, gy0, din, 2);
        in_link[(size_t)J(i, j) * 4 + din] = k2;
      } } }
  const int n_ext_links = (int)link_i.size() / 8;
  auto lane_pt = [&](int link, int num, bool end, double& px, double& py) { const double* g = &link_geom[4 * (size_t)link]; int d = link_dir[link]; double off = (num + 0.5) * lane_w;
    double bx = end ? g[2] : g[0], by = end ? g[3] : g[1]; px = bx + DY[d] * off; py = by - DX[d] * off; };
  struct Move { int node, a, o, din; };
  std::vector<Move> moves; std::vector<int> exit_in_count(n_ext_links, 0);
  const int turns[3] = {0, 3, 1};
  for (int i = 0; i < nx; i++) for (int j = 0; j < ny; j++) for (int din = 0; din < 4; din++) { int a = in_link[(size_t)J(i, j) * 4 + din]; if (a < 0) continue;
    for (int t = 0; t < 3; t++) { int dout = (din + turns[t]) % 4; int o = out_link[(size_t)J(i, j) * 4 + dout]; if (o < 0) continue; moves.push_back({J(i, j), a, o, din}); exit_in_count[o] += NL; } }
  const int n_ext_lanes = (int)lane_i.size() / 4;
  std::vector<std::vector<int>> succ_rows((size_t)n_ext_lanes);       // per external lane: (key, next_lane) pairs
  std::vector<int> int_succ_key, int_succ_next;                       // internal lanes in creation order
  std::vector<std::vector<int>> sl_rows((size_t)n_ext_lanes);          // (ctrl, up_link, up_num, down_link, down_num)
  std::vector<std::vector<int>> sig_in((size_t)nx * ny * 2), sig_out((size_t)nx * ny * 2);
  for (auto& m : moves) {
    int seq = (m.din == 0 || m.din == 2) ? 0 : 1; sig_in[(size_t)m.node * 2 + seq].push_back(m.a); sig_out[(size_t)m.node * 2 + seq].push_back(m.o);
    for (int k = 0; k < NL; k++) { double ax, ay, ox, oy; lane_pt(m.a, k, true, ax, ay); lane_pt(m.o, k, false, ox, oy); double len = std::hypot(ox - ax, oy - ay);
      int ent = link_i[8 * (size_t)m.a] + k;
      int il = add_link(m.node, m.node, 1, len, exit_in_count[m.o] > 1 ? 'C' : 'R', m.node, ent, ax, ay, ox, oy, m.din, 3);
      int ilane = link_i[8 * (size_t)il];
      succ_rows[ent].push_back(m.o); succ_rows[ent].push_back(ilane);
      int_succ_key.push_back(m.o); int_succ_next.push_back(link_i[8 * (size_t)m.o] + k);
      sl_rows[ent].insert(sl_rows[ent].end(), {m.node, m.a, k, m.o, k}); } }
  const int n_lanes = (int)lane_i.size() / 4;
  std::vector<int> succ_off(n_lanes + 1, 0), succ_i, sl_off(n_lanes + 1, 0), sl_i; std::vector<double> succ_f, sl_f;
  for (int l = 0; l < n_lanes; l++) {
    if (l < n_ext_lanes) { for (size_t q = 0; q + 1 < succ_rows[l].size(); q += 2) { succ_i.insert(succ_i.end(), {succ_rows[l][q], succ_rows[l][q + 1], 1}); succ_f.push_back(1.0); }
      for (size_t q = 0; q + 4 < sl_rows[l].size(); q += 5) { sl_i.insert(sl_i.end(), {sl_rows[l][q], sl_rows[l][q + 1], sl_rows[l][q + 2], sl_rows[l][q + 3], sl_rows[l][q + 4], 0}); sl_f.push_back(lane_f[l]); } }
    else { int q = l - n_ext_lanes; succ_i.insert(succ_i.end(), {int_succ_key[q], int_succ_next[q], 1}); succ_f.push_back(1.0); }
    succ_off[l + 1] = (int)succ_i.size() / 3; sl_off[l + 1] = (int)sl_i.size() / 6; }
  std::vector<int> ctrl_i, seq_i, sig_i; std::vector<double> ctrl_f, seq_f, sig_f;
  for (int c = 0; c < nx * ny; c++) { ctrl_i.push_back((int)seq_i.size() / 2); ctrl_i.push_back(2); ctrl_f.push_back(2 * (green + amber));
    for (int q = 0; q < 2; q++) { auto& a = sig_in[(size_t)c * 2 + q]; auto& o = sig_out[(size_t)c * 2 + q]; seq_i.push_back((int)sig_i.size() / 2); seq_i.push_back((int)a.size()); seq_f.push_back(green + amber);
      for (size_t k = 0; k < a.size(); k++) { sig_i.push_back(a[k]); sig_i.push_back(o[k]); sig_f.insert(sig_f.end(), {green, amber, 0.0}); } } }
  // routes: two staircase itineraries per junction->junction link, at most route_len links
  std::vector<int> link_at_i(n_ext_links, -1), link_at_j(n_ext_links, -1);
  for (int i = 0; i < nx; i++) for (int j = 0; j < ny; j++) for (int d = 0; d < 4; d++) { int k = out_link[(size_t)J(i, j) * 4 + d]; if (k >= 0) { link_at_i[k] = i; link_at_j[k] = j; } }
  std::vector<int> route_off(1, 0), route_links; std::vector<int> route_of_link((size_t)n_ext_links * 2, -1);
  Rng rr((uint64_t)seed * 7919 + 13);
  for (int k = 0; k < n_ext_links; k++) { if (link_kind[k] != 0) continue; int d = link_dir[k];
    for (int side = 0; side < 2; side++) { int d2 = (d + (side == 0 ? 1 : 3)) % 4; int ci = link_at_i[k] + DX[d], cj = link_at_j[k] + DY[d];
      route_links.push_back(k); int len = 1;
      while (len < route_len) { int c0 = out_link[(size_t)J(ci, cj) * 4 + d], c1 = out_link[(size_t)J(ci, cj) * 4 + d2]; int pick;
        if (c0 >= 0 && c1 >= 0) pick = (rr.next() & 1) ? d2 : d; else pick = c0 >= 0 ? d : d2;
        int o = out_link[(size_t)J(ci, cj) * 4 + pick]; route_links.push_back(o); len++;
        if (link_kind[o] == 1) break;
        ci += DX[pick]; cj += DY[pick]; }
      route_of_link[(size_t)k * 2 + side] = (int)route_off.size() - 1; route_off.push_back((int)route_links.size()); } }
  // vehicles: counts for every seeded lane (global ids), rows only for the lanes this rank owns
  const double inv_kx = 1.0 / kx, vmax = std::min(vx, vit_reg);
  const double lo = -block, hi = nx * block;
  std::vector<int> iv_i, iv_id; std::vector<double> iv_f; long long next_id = 0;
  for (int k = 0; k < n_ext_links; k++) { if (link_kind[k] != 0) continue;
    int down = link_i[8 * (size_t)k + 3]; double xdown = node_xy[2 * (size_t)down];
    int owner = world <= 1 ? 0 : std::min(world - 1, std::max(0, (int)std::floor((xdown - lo) / (hi - lo) * world)));
    for (int num = 0; num < NL; num++) { int lane = link_i[8 * (size_t)k] + num; Rng r((uint64_t)seed * 1000003ull + (uint64_t)lane);
      int m = (int)per_lane + (r.uni() < per_lane - (int)per_lane ? 1 : 0); m = std::min(m, (int)(L / inv_kx) - 1); if (m <= 0) continue;
      if (owner == rank) { double slot = L / m, jit = std::max(0.0, slot - inv_kx - 1e-6); std::vector<double> pos(m);
        for (int q = 0; q < m; q++) pos[q] = (q + 0.5) * slot - 0.5 * jit + r.uni() * jit * 0.999;
        std::sort(pos.begin(), pos.end());
        for (int q = 0; q < m; q++) { double s = q + 1 < m ? pos[q + 1] - pos[q] : 1e9; double v = std::min(vmax, std::max(0.0, -w * (s * kx - 1.0)));
          int rt = route_of_link[(size_t)k * 2 + (int)(r.next() & 1)];
          iv_i.insert(iv_i.end(), {0, lane, rt, 0}); iv_f.insert(iv_f.end(), {pos[q], v, 0.0}); iv_id.push_back((int)(next_id + q)); } }
      next_id += m; } }
  if (n_vehicles_total) *n_vehicles_total = next_id;
  std::vector<double> params(32, 0.0); params[0] = 1.0; params[1] = n_steps; params[2] = seed; params[4] = 1.0; params[6] = 50.0;
  std::vector<double> cls(24, 0.0); cls[0] = w; cls[1] = kx; cls[2] = vx; cls[3] = 1.0; cls[5] = 4.0; cls[8] = 1; cls[9] = 1.5; cls[10] = 1e9;
  Writer W(path); if (!W.f) return -1;
  W.D("params", params); W.D("cls", cls); W.I("link_i", link_i); W.D("link_f", link_f); W.D("link_geom", link_geom); W.I("lane_i", lane_i); W.D("lane_f", lane_f);
  W.I("succ_off", succ_off); W.I("succ_i", succ_i); W.D("succ_f", succ_f); W.I("sl_off", sl_off); W.I("sl_i", sl_i); W.D("sl_f", sl_f);
  W.I("ctrl_i", ctrl_i); W.D("ctrl_f", ctrl_f); W.I("seq_i", seq_i); W.D("seq_f", seq_f); W.I("sig_i", sig_i); W.D("sig_f", sig_f);
  W.I("node_i", node_i); W.D("node_xy", node_xy); W.I("route_off", route_off); W.I("route_links", route_links);
  W.I("exit_i", exit_i); W.D("exit_f", exit_f); W.I("iv_i", iv_i); W.D("iv_f", iv_f); W.I("iv_id", iv_id);
  const char cn[] = "VL"; W.sec("names_cls", 2, cn, 3);
  return W.close() ? (int)(iv_id.size()) : -1;
}
