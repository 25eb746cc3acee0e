// flatnet.h — host-side flattened network tables (SYMFLAT1, include/symflat.h).
// Replaces the object graph Reseau::LoadReseauXML builds (symuvia/src/reseau.cpp:4931-9612) by
// dense arrays that can be uploaded to HBM as they are.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/symflat.h"

namespace symb200 {

struct FlatNet {
  std::vector<double> params, cls, link_f, lane_f, succ_f, sl_f, ctrl_f, seq_f, sig_f, dem_f, dest_f, droute_f,
      lanecoef_f, typecoef_f, exit_f, iv_f, link_geom, node_xy;
  std::vector<int> link_i, lane_i, succ_off, succ_i, sl_off, sl_i, ctrl_i, seq_i, sig_i, node_i, route_off, route_links,
      entry_i, dest_i, droute_i, exit_i, iv_i, iv_id,
      cvg_i, cvgam_i, pt_i, ptam_i, cafmvt_i, cafmvt_links, lcam_off, lcam_links, lcav_off, lcav_links;   // merge points (row A10)
  std::vector<double> cvg_f, cvg_tf;
  std::vector<std::string> link_names, node_names, cls_names;
  std::string error;

  int n_links() const { return (int)link_i.size() / 8; }
  int n_lanes() const { return (int)lane_i.size() / 4; }
  int n_cls() const { return (int)cls.size() / SYMFLAT_CLS_STRIDE; }
  int n_entries() const { return (int)entry_i.size() / 8; }
  int n_exits() const { return (int)exit_i.size() / 2; }
  int n_init() const { return (int)iv_i.size() / 4; }
  int n_ctrl() const { return (int)ctrl_i.size() / 2; }
  int n_sl() const { return (int)sl_i.size() / 6; }
  int n_cvg() const { return (int)cvg_i.size() / 8; }
  int n_pt() const { return (int)pt_i.size() / 4; }
  int n_nodes() const { return (int)node_i.size() / 4; }

  static void split_names(const std::vector<unsigned char>& raw, std::vector<std::string>& out) {
    std::string cur;
    for (unsigned char c : raw) { if (c == 0) { out.push_back(cur); cur.clear(); } else cur.push_back((char)c); }
  }

  bool load(const char* path) {
    FILE* f = fopen(path, "rb");
    if (!f) { error = std::string("cannot open ") + path; return false; }
    char magic[8]; uint32_t n = 0;
    if (fread(magic, 1, 8, f) != 8 || memcmp(magic, SYMFLAT_MAGIC, 8) != 0 || fread(&n, 4, 1, f) != 1) {
      fclose(f); error = "bad SYMFLAT1 header"; return false; }
    std::map<std::string, std::vector<int>*> im = {
        {"link_i", &link_i}, {"lane_i", &lane_i}, {"succ_off", &succ_off}, {"succ_i", &succ_i}, {"sl_off", &sl_off},
        {"sl_i", &sl_i}, {"ctrl_i", &ctrl_i}, {"seq_i", &seq_i}, {"sig_i", &sig_i}, {"node_i", &node_i},
        {"route_off", &route_off}, {"route_links", &route_links}, {"entry_i", &entry_i}, {"dest_i", &dest_i},
        {"droute_i", &droute_i}, {"exit_i", &exit_i}, {"iv_i", &iv_i}, {"iv_id", &iv_id},
        {"cvg_i", &cvg_i}, {"cvgam_i", &cvgam_i}, {"pt_i", &pt_i}, {"ptam_i", &ptam_i}, {"cafmvt_i", &cafmvt_i}, {"cafmvt_links", &cafmvt_links},
        {"lcam_off", &lcam_off}, {"lcam_links", &lcam_links}, {"lcav_off", &lcav_off}, {"lcav_links", &lcav_links}};
    std::map<std::string, std::vector<double>*> dm = {
        {"params", &params}, {"cls", &cls}, {"link_f", &link_f}, {"lane_f", &lane_f}, {"succ_f", &succ_f},
        {"sl_f", &sl_f}, {"ctrl_f", &ctrl_f}, {"seq_f", &seq_f}, {"sig_f", &sig_f}, {"dem_f", &dem_f},
        {"dest_f", &dest_f}, {"droute_f", &droute_f}, {"lanecoef_f", &lanecoef_f}, {"typecoef_f", &typecoef_f},
        {"exit_f", &exit_f}, {"iv_f", &iv_f}, {"link_geom", &link_geom}, {"cvg_f", &cvg_f}, {"cvg_tf", &cvg_tf}, {"node_xy", &node_xy}};
    for (uint32_t s = 0; s < n; s++) {
      char name[25] = {0}; uint32_t dt; uint64_t cnt;
      if (fread(name, 1, 24, f) != 24 || fread(&dt, 4, 1, f) != 1 || fread(&cnt, 8, 1, f) != 1) { fclose(f); error = "truncated section header"; return false; }
      size_t isz = dt == 0 ? 4 : dt == 1 ? 8 : 1, bytes = cnt * isz;
      std::string nm(name);
      bool ok = true;
      if (dt == 0 && im.count(nm)) { im[nm]->resize(cnt); ok = !cnt || fread(im[nm]->data(), 1, bytes, f) == bytes; }
      else if (dt == 1 && dm.count(nm)) { dm[nm]->resize(cnt); ok = !cnt || fread(dm[nm]->data(), 1, bytes, f) == bytes; }
      else {
        std::vector<unsigned char> raw(bytes);
        ok = !bytes || fread(raw.data(), 1, bytes, f) == bytes;
        if (nm == "names_link") split_names(raw, link_names);
        else if (nm == "names_node") split_names(raw, node_names);
        else if (nm == "names_cls") split_names(raw, cls_names);
      }
      if (!ok) { fclose(f); error = "truncated section " + nm; return false; }
      size_t pad = (8 - bytes % 8) % 8; char tmp[8];
      if (pad && fread(tmp, 1, pad, f) != pad) { fclose(f); error = "truncated padding"; return false; }
    }
    fclose(f);
    return validate();
  }

  // writes the tables as a SYMFLAT1 file (include/symflat.h)
  bool save(const char* path) const {
    FILE* f = fopen(path, "wb"); if (!f) return false;
    std::vector<std::pair<std::string, std::pair<int, std::pair<const void*, size_t>>>> secs;
    auto I = [&](const char* n, const std::vector<int>& v) { secs.push_back({n, {0, {v.data(), v.size()}}}); };
    auto D = [&](const char* n, const std::vector<double>& v) { secs.push_back({n, {1, {v.data(), v.size()}}}); };
    auto join = [](const std::vector<std::string>& v) { std::string s; for (auto& x : v) { s += x; s.push_back('\0'); } return s; };
    const std::string nl = join(link_names), nn = join(node_names), nc = join(cls_names);
    D("params", params); D("cls", cls); I("link_i", link_i); D("link_f", link_f); D("link_geom", link_geom); I("lane_i", lane_i); D("lane_f", lane_f);
    I("succ_off", succ_off); I("succ_i", succ_i); D("succ_f", succ_f); I("sl_off", sl_off); I("sl_i", sl_i); D("sl_f", sl_f);
    I("ctrl_i", ctrl_i); D("ctrl_f", ctrl_f); I("seq_i", seq_i); D("seq_f", seq_f); I("sig_i", sig_i); D("sig_f", sig_f);
    I("node_i", node_i); D("node_xy", node_xy); I("route_off", route_off); I("route_links", route_links);
    I("entry_i", entry_i); D("dem_f", dem_f); I("dest_i", dest_i); D("dest_f", dest_f); I("droute_i", droute_i); D("droute_f", droute_f);
    D("lanecoef_f", lanecoef_f); D("typecoef_f", typecoef_f); I("exit_i", exit_i); D("exit_f", exit_f); I("iv_i", iv_i); D("iv_f", iv_f);
    if (!iv_id.empty()) I("iv_id", iv_id);
    if (!cvg_i.empty() || !lcam_links.empty() || !lcav_links.empty()) { I("cvg_i", cvg_i); D("cvg_f", cvg_f); D("cvg_tf", cvg_tf); I("cvgam_i", cvgam_i); I("pt_i", pt_i); I("ptam_i", ptam_i);
      I("cafmvt_i", cafmvt_i); I("cafmvt_links", cafmvt_links); I("lcam_off", lcam_off); I("lcam_links", lcam_links); I("lcav_off", lcav_off); I("lcav_links", lcav_links); }
    secs.push_back({"names_link", {2, {nl.data(), nl.size()}}}); secs.push_back({"names_node", {2, {nn.data(), nn.size()}}}); secs.push_back({"names_cls", {2, {nc.data(), nc.size()}}});
    fwrite(SYMFLAT_MAGIC, 1, 8, f); uint32_t n = (uint32_t)secs.size(); fwrite(&n, 4, 1, f);
    for (auto& s : secs) { char nm[24] = {0}; strncpy(nm, s.first.c_str(), 23); uint32_t dt = (uint32_t)s.second.first; uint64_t cnt = s.second.second.second;
      fwrite(nm, 1, 24, f); fwrite(&dt, 4, 1, f); fwrite(&cnt, 8, 1, f);
      const size_t bytes = cnt * (dt == 0 ? 4 : dt == 1 ? 8 : 1); if (bytes) fwrite(s.second.second.first, 1, bytes, f);
      const size_t pad = (8 - bytes % 8) % 8; const char z[8] = {0}; if (pad) fwrite(z, 1, pad, f); }
    return fclose(f) == 0;
  }

  bool validate() {
    if (params.size() < SYMFLAT_P_COUNT) { error = "params missing"; return false; }
    if (cls.empty() || cls.size() % SYMFLAT_CLS_STRIDE) { error = "cls table malformed"; return false; }
    if (link_i.size() % 8 || link_f.size() != (size_t)n_links() * 2) { error = "link tables malformed"; return false; }
    if (lane_i.size() % 4 || lane_f.size() != (size_t)n_lanes()) { error = "lane tables malformed"; return false; }
    if (succ_off.size() != (size_t)n_lanes() + 1 || sl_off.size() != (size_t)n_lanes() + 1) { error = "CSR offsets malformed"; return false; }
    if (route_off.empty()) route_off.push_back(0);
    if (lcam_off.size() != (size_t)n_links() + 1) { lcam_off.assign(n_links() + 1, 0); lcam_links.clear(); }
    if (lcav_off.size() != (size_t)n_links() + 1) { lcav_off.assign(n_links() + 1, 0); lcav_links.clear(); }
    if (cvg_f.size() != (size_t)n_cvg() * 4 || cvg_tf.size() != (size_t)n_cvg() * n_cls()) { error = "merge tables malformed"; return false; }
    return true;
  }
};

}  // namespace symb200
