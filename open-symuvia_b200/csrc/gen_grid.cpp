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

extern "C" int symgen_grid2(const char* path, int nx, int ny, double per_lane, int seed, int route_len, int n_steps, int rank, int world,
                            long long* n_vehicles_total, int merges) {
  const double block = 200.0, setback = 12.0, lane_w = 3.0, vit_reg = 13.89, green = 30.0, amber = 3.0;
  const int NL = 2;
  const double w = -5.8823, kx = 0.17, vx = 25.0;
  const double L = block - 2 * setback;
  auto J = [&](int i, int j) { return i * ny + j; };
  std::vector<int> node_i; std::vector<double> node_xy;
  for (int i = 0; i < nx; i++) for (int j = 0; j < ny; j++) { node_i.insert(node_i.end(), {(int)'C', J(i, j), 0, 0}); node_xy.push_back(i * block); node_xy.push_back(j * block); }
  std::vector<int> link_i, lane_i; std::vector<double> link_f, link_geom, lane_f;
  std::vector<int> link_dir, link_kind;                      // kind: 0 'L' junction->junction, 1 exit stub, 2 entry stub, 3 internal
  std::vector<std::string> link_name;                        // labels as in synth.py: they fix the creation order inside a junction and the order of the merges
  const char DN[5] = "ENWS";
  auto nm = [&](char pre, int i, int j, int d) { return std::string(1, pre) + std::to_string(i) + "_" + std::to_string(j) + std::string(1, DN[d]); };
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
        int k = add_link(J(i, j), J(i2, j2), NL, L, 'D', -1, -1, gx0, gy0, i2 * block - DX[d] * setback, j2 * block - DY[d] * setback, d, 0); link_name.push_back(nm('L', i, j, d));
        out_link[(size_t)J(i, j) * 4 + d] = k; in_link[(size_t)J(i2, j2) * 4 + d] = k;
      } else {
        int s = (int)node_i.size() / 4; node_i.insert(node_i.end(), {(int)'S', -1, 0, 0}); node_xy.push_back(x0 + DX[d] * block); node_xy.push_back(y0 + DY[d] * block);
        int k = add_link(J(i, j), s, NL, L, 'S', -1, -1, gx0, gy0, x0 + DX[d] * (block - setback), y0 + DY[d] * (block - setback), d, 1); link_name.push_back(nm('X', i, j, d));
        out_link[(size_t)J(i, j) * 4 + d] = k; exit_i.push_back(s); exit_i.push_back(k); exit_f.push_back(-1.0);
        int e = (int)node_i.size() / 4; node_i.insert(node_i.end(), {(int)'E', -1, 0, 0}); node_xy.push_back(x0 + DX[d] * block); node_xy.push_back(y0 + DY[d] * block);
        int din = (d + 2) % 4;
        int k2 = add_link(e, J(i, j), NL, L, 'D', -1, -1, x0 + DX[d] * (block - setback), y0 + DY[d] * (block - setback), gx0, gy0, din, 2); link_name.push_back(nm('N', i, j, d));
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
  for (auto& m : moves) { int seq = (m.din == 0 || m.din == 2) ? 0 : 1; sig_in[(size_t)m.node * 2 + seq].push_back(m.a); sig_out[(size_t)m.node * 2 + seq].push_back(m.o); }
  // internal links in the creation order of CarrefourAFeuxEx::Init (CarrefourAFeuxEx.cpp:331-480): entry-lane label, exit-link label, exit lane;
  // one Convergent per exit link fed by several internal links (:489-919), reference priority levels from the geometry (:1129-1166)
  struct Cv { std::string name; int node, down; std::vector<int> am; std::vector<std::vector<std::pair<int, int>>> vam; };
  std::vector<Cv> cvs; std::vector<int> mv_i, mv_l; std::vector<std::vector<int>> lcam((size_t)n_ext_links), lcav((size_t)n_ext_links);
  std::vector<std::vector<int>> int_cam;                                // upstream link of every internal link (its entry link)
  { size_t m0 = 0;
    while (m0 < moves.size()) { size_t m1 = m0; while (m1 < moves.size() && moves[m1].node == moves[m0].node) m1++;
      struct E { std::string ka, ko; int k, a, o, din; };
      std::vector<E> es;
      for (size_t q = m0; q < m1; q++) for (int k = 0; k < NL; k++) es.push_back({link_name[moves[q].a] + "V" + std::to_string(k + 1), link_name[moves[q].o], k, moves[q].a, moves[q].o, moves[q].din});
      std::sort(es.begin(), es.end(), [](const E& x, const E& y) { if (x.ka != y.ka) return x.ka < y.ka; if (x.ko != y.ko) return x.ko < y.ko; return x.k < y.k; });
      std::vector<int> exits; std::vector<std::vector<std::pair<int, int>>> feeds;
      const int node = moves[m0].node;
      for (auto& e : es) { double ax, ay, ox, oy; lane_pt(e.a, e.k, true, ax, ay); lane_pt(e.o, e.k, false, ox, oy); double len = std::hypot(ox - ax, oy - ay);
        int ent = link_i[8 * (size_t)e.a] + e.k;
        int il = add_link(node, node, 1, len, exit_in_count[e.o] > 1 ? 'C' : 'R', node, ent, ax, ay, ox, oy, e.din, 3);
        int ilane = link_i[8 * (size_t)il];
        succ_rows[ent].push_back(e.o); succ_rows[ent].push_back(ilane);
        int_succ_key.push_back(e.o); int_succ_next.push_back(link_i[8 * (size_t)e.o] + e.k);
        sl_rows[ent].insert(sl_rows[ent].end(), {node, e.a, e.k, e.o, e.k});
        size_t xi = 0; while (xi < exits.size() && exits[xi] != e.o) xi++;
        if (xi == exits.size()) { exits.push_back(e.o); feeds.emplace_back(); }
        feeds[xi].push_back({il, e.k});
        mv_i.insert(mv_i.end(), {node, e.a, e.o, (int)mv_l.size(), 1}); mv_l.push_back(il);
        int_cam.push_back({e.a}); lcav[e.a].push_back(il); }
      for (size_t xi = 0; xi < exits.size(); xi++) { const int o = exits[xi]; for (auto& fd : feeds[xi]) lcam[o].push_back(fd.first);
        if (feeds[xi].size() < 2) continue;
        Cv cv; cv.name = "J" + std::to_string(node / ny) + "_" + std::to_string(node % ny) + "_C0_" + link_name[o]; cv.node = node; cv.down = o;
        for (auto& fd : feeds[xi]) cv.am.push_back(fd.first);
        for (int kk = 0; kk < NL; kk++) { std::vector<std::pair<double, int>> ang; double vx0, vy0, vx1, vy1; lane_pt(o, kk, false, vx0, vy0); lane_pt(o, kk, true, vx1, vy1);
          std::vector<int> lanes_kk;
          for (auto& fd : feeds[xi]) if (fd.second == kk) { const int ln = link_i[8 * (size_t)fd.first]; lanes_kk.push_back(ln); const double* g = &link_geom[4 * (size_t)fd.first];
            double a1 = vx1 - vx0, a2 = vy1 - vy0, b1 = g[0] - vx0, b2 = g[1] - vy0; double t = std::atan2(a1 * b2 - a2 * b1, a1 * b1 + a2 * b2); if (t < 0) t += 2 * 3.14159265358979323846;
            ang.push_back({t, ln}); }
          std::sort(ang.begin(), ang.end());
          std::vector<std::pair<int, int>> v; for (int ln : lanes_kk) { int lev = 0; for (size_t q = 0; q < ang.size(); q++) if (ang[q].second == ln) lev = (int)ang.size() - (int)q; v.push_back({ln, lev}); }
          cv.vam.push_back(v); }
        cvs.push_back(std::move(cv)); }
      m0 = m1; } }
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
  if (merges) {
    std::sort(cvs.begin(), cvs.end(), [](const Cv& x, const Cv& y) { return x.name < y.name; });      // Reseau::Liste_convergents: std::map by id
    std::vector<int> cvg_i, cvgam, pt_i, ptam; std::vector<double> cvg_f, cvg_tf;
    for (size_t ci = 0; ci < cvs.size(); ci++) { const Cv& c = cvs[ci];
      cvg_i.insert(cvg_i.end(), {c.node, c.down, (int)pt_i.size() / 4, (int)c.vam.size(), (int)cvgam.size(), (int)c.am.size(), 0, 0});
      cvg_f.insert(cvg_f.end(), {30.0, 1.0, 1.0, 5.0}); cvg_tf.push_back(0.0); cvgam.insert(cvgam.end(), c.am.begin(), c.am.end());
      for (size_t kk = 0; kk < c.vam.size(); kk++) { pt_i.insert(pt_i.end(), {(int)ci, link_i[8 * (size_t)c.down] + (int)kk, (int)ptam.size() / 2, (int)c.vam[kk].size()});
        for (auto& a : c.vam[kk]) { ptam.push_back(a.first); ptam.push_back(a.second); } } }
    const int K = (int)link_i.size() / 8; std::vector<int> cam_off(K + 1, 0), cam_l, cav_off(K + 1, 0), cav_l;
    for (int k = 0; k < K; k++) { const std::vector<int>& a = k < n_ext_links ? lcam[k] : int_cam[k - n_ext_links]; cam_l.insert(cam_l.end(), a.begin(), a.end()); cam_off[k + 1] = (int)cam_l.size();
      if (k < n_ext_links) cav_l.insert(cav_l.end(), lcav[k].begin(), lcav[k].end()); else cav_l.push_back(int_succ_key[k - n_ext_links]);
      cav_off[k + 1] = (int)cav_l.size(); }
    W.I("cvg_i", cvg_i); W.D("cvg_f", cvg_f); W.D("cvg_tf", cvg_tf); W.I("cvgam_i", cvgam); W.I("pt_i", pt_i); W.I("ptam_i", ptam); W.I("cafmvt_i", mv_i); W.I("cafmvt_links", mv_l);
    W.I("lcam_off", cam_off); W.I("lcam_links", cam_l); W.I("lcav_off", cav_off); W.I("lcav_links", cav_l);
  }
  const char cn[] = "VL"; W.sec("names_cls", 2, cn, 3);
  return W.close() ? (int)(iv_id.size()) : -1;
}
extern "C" int symgen_grid(const char* path, int nx, int ny, double per_lane, int seed, int route_len, int n_steps, int rank, int world, long long* n_vehicles_total) {
  return symgen_grid2(path, nx, ny, per_lane, seed, route_len, n_steps, rank, world, n_vehicles_total, 1);
}
