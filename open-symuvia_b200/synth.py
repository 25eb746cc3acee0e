"""Synthetic networks + demand for the BASELINE.json configs, emitted as SYMFLAT1 tables
(and, for small cases, as a reference-format `ROOT_SYMUBRUIT` XML document so a user with
a real SymuVia build can cross-check: SURVEY.md Appendix B).

What is flattened here is what the reference builds at load time:
  * links / lanes (`TuyauMicro`, `VoieMicro`: `symuvia/src/tuyau.cpp:1503-1615`);
  * signalised junctions (`CarrefourAFeuxEx::Init`, `symuvia/src/CarrefourAFeuxEx.cpp:215-1196`):
    one single-lane internal link per allowed movement (entry lane -> exit lane), running
    straight from the entry lane's downstream end to the exit lane's upstream end
    (`:660-732`), entry links typed 'D' (`:379`), internal links 'C' (merge) or 'R' when the
    exit has a single upstream movement (`:860-880`);
  * stop lines (`CarrefourAFeuxEx::FinInit`, `:1205-1275`): `ligne_feu = 0` puts one
    `LigneDeFeu` per (entry lane, exit link, exit lane) at the entry lane's end;
  * fixed-time signal plans (`ControleurDeFeux`, `symuvia/src/ControleurDeFeux.cpp:203-307`).
All generators are deterministic (NumPy `default_rng(seed)`).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import flat as F

VL = dict(name="VL", w=-5.8823, kx=0.17, vx=25.0, esp=1.0, decel=0.0, tau_lc=4.0, pi_rab=0.0,
          law=F.LAW_NEWELL, plages=[(1.5, 1e9)])
PL = dict(name="PL", w=-4.5, kx=0.08, vx=22.0, esp=1.0, decel=0.0, tau_lc=4.0, pi_rab=0.0,
          law=F.LAW_NEWELL, plages=[(1.0, 1e9)])


class NetBuilder:
    """Accumulates the flat tables. Ids are dense ints in creation order."""

    def __init__(self, dt=1.0, n_steps=100, seed=42, accbornee=False, relax=1.0, chgtvoie=False,
                 dst_chgtvoie=50.0):
        self.params = np.zeros(F.P_NPARAMS)
        self.params[F.P_DT] = dt
        self.params[F.P_NSTEPS] = n_steps
        self.params[F.P_SEED] = seed
        self.params[F.P_ACCBORNEE] = float(accbornee)
        self.params[F.P_RELAX] = relax
        self.params[F.P_CHGTVOIE] = float(chgtvoie)
        self.params[F.P_DSTCHGT] = dst_chgtvoie
        self.classes: List[dict] = []
        self.nodes: List[List[int]] = []          # type, ctrl, 0, 0
        self.node_names: List[str] = []
        self.node_xy: List[Tuple[float, float]] = []
        self.links: List[List[int]] = []
        self.link_f: List[List[float]] = []
        self.link_names: List[str] = []
        self.link_geom: List[Tuple[float, float, float, float]] = []
        self.link_lane_w: List[float] = []
        self.lanes: List[List[int]] = []
        self.lane_f: List[float] = []
        self.succ: List[List[Tuple[int, int, int, float]]] = []     # per lane
        self.sl: List[List[Tuple[int, int, int, int, int, float]]] = []
        self.ctrl: List[dict] = []
        self.routes: List[List[int]] = []
        self.entries: List[dict] = []
        self.exits: List[Tuple[int, int, float]] = []
        self.iv_i: List[List[int]] = []
        self.iv_f: List[List[float]] = []
        # merge points (row A10): Convergent objects in Reseau::Liste_convergents order (std::map by id, reseau.h:201)
        self.cvgs: List[dict] = []
        self.caf_mvts: List[Tuple[int, int, int, List[int]]] = []   # CarrefourAFeuxEx::m_LstMouvements: node, entry link, exit link, internal links
        self.link_cam: Dict[int, List[int]] = {}   # link -> m_LstTuyAm of its upstream connection (reference order)
        self.link_cav: Dict[int, List[int]] = {}   # link -> m_LstTuyAv of its downstream connection

    # -- topology ---------------------------------------------------------------------------
    def add_class(self, c: dict) -> int:
        self.classes.append(c)
        return len(self.classes) - 1

    def add_node(self, typ: str, name: str, xy=(0.0, 0.0), ctrl=-1) -> int:
        self.nodes.append([ord(typ), ctrl, 0, 0])
        self.node_names.append(name)
        self.node_xy.append(xy)
        return len(self.nodes) - 1

    def add_link(self, name: str, up: int, down: int, n_lanes: int, length: float, vit_reg: float,
                 type_aval: str, brick: int = -1, prev_lane: int = -1, geom=(0.0, 0.0, 0.0, 0.0),
                 lane_lengths: Optional[List[float]] = None, lane_w: float = 3.0) -> int:
        k = len(self.links)
        self.link_lane_w.append(lane_w)
        first = len(self.lanes)
        self.links.append([first, n_lanes, up, down, ord(type_aval), brick, 0, 0])
        self.link_f.append([length, vit_reg])
        self.link_names.append(name)
        self.link_geom.append(geom)
        for n in range(n_lanes):
            self.lanes.append([k, n, prev_lane, 0])
            self.lane_f.append(length if lane_lengths is None else lane_lengths[n])
            self.succ.append([])
            self.sl.append([])
        return k

    def lane_of(self, link: int, num: int) -> int:
        return self.links[link][0] + num

    def add_succ(self, lane: int, key_link: int, next_lane: int, rate=1.0, draws=1):
        self.succ[lane].append((key_link, next_lane, draws, rate))

    def add_stopline(self, lane: int, pos: float, ctrl: int, up_link: int, up_num: int,
                     down_link: int, down_num: int):
        self.sl[lane].append((ctrl, up_link, up_num, down_link, down_num, pos))

    def add_ctrl(self, cycle: float, seqs: List[dict]) -> int:
        """seqs: [{len: float, signals: [(link_in, link_out, green, amber, delay)]}]"""
        self.ctrl.append(dict(cycle=cycle, seqs=seqs))
        return len(self.ctrl) - 1

    def add_route(self, links: List[int]) -> int:
        self.routes.append(list(links))
        return len(self.routes) - 1

    def add_entry(self, node: int, link: int, demand: List[Tuple[float, float]],
                  dests: List[Tuple[int, float, List[Tuple[int, float]]]],
                  lane_coefs: Optional[List[float]] = None, type_coefs: Optional[List[float]] = None,
                  exponential=False) -> int:
        """demand: [(duration_s, level veh/s)]; dests: [(exit_idx, coeff, [(route, coeff)])]."""
        self.entries.append(dict(node=node, link=link, demand=demand, dests=dests, lane_coefs=lane_coefs,
                                 type_coefs=type_coefs, exponential=exponential))
        return len(self.entries) - 1

    def add_exit(self, node: int, link: int, capacity=-1.0) -> int:
        self.exits.append((node, link, capacity))
        return len(self.exits) - 1

    def add_convergent(self, name: str, node: int, down_link: int, am_links: List[int], points: List[dict], tagreg=30.0, gamma=1.0,
                       mu=1.0, pos_cpt_av=5.0, tf: Optional[List[float]] = None) -> dict:
        """One `Convergent` (convergent.cpp:159-255): `points` = one dict per lane of the downstream link,
        {down_lane, vam: [(upstream lane, reference priority level)]} in `m_LstVAm` order; `am_links` = `m_LstTuyAm` order
        (one ponctual sensor each at length - 5, FinInit :230-255); sensor of the downstream link at `pos_cpt_av`."""
        c = dict(name=name, node=node, down_link=down_link, am_links=list(am_links), points=points, tagreg=tagreg, gamma=gamma, mu=mu,
                 pos_cpt_av=pos_cpt_av, tf=tf)
        self.cvgs.append(c)
        return c

    def add_init_vehicle(self, cls: int, lane: int, route: int, route_pos: int, pos: float, v: float,
                         a: float = 0.0):
        self.iv_i.append([cls, lane, route, route_pos])
        self.iv_f.append([pos, v, a])

    # -- emit -------------------------------------------------------------------------------
    def sections(self) -> Dict[str, np.ndarray]:
        T = len(self.classes)
        cls = np.zeros((T, F.CLS_STRIDE))
        for i, c in enumerate(self.classes):
            cls[i, F.C_W], cls[i, F.C_KX], cls[i, F.C_VX] = c["w"], c["kx"], c["vx"]
            cls[i, F.C_ESP], cls[i, F.C_DECEL] = c["esp"], c["decel"]
            cls[i, F.C_TAULC], cls[i, F.C_PIRAB], cls[i, F.C_LAW] = c["tau_lc"], c["pi_rab"], c["law"]
            cls[i, F.C_NPLAGE] = len(c["plages"])
            for j, (ax, vs) in enumerate(c["plages"][:4]):
                cls[i, F.C_PLAGE0 + 2 * j] = ax
                cls[i, F.C_PLAGE0 + 2 * j + 1] = vs
        L = len(self.lanes)
        succ_off = np.zeros(L + 1, np.int32)
        sl_off = np.zeros(L + 1, np.int32)
        succ_i, succ_f, sl_i, sl_f = [], [], [], []
        for l in range(L):
            for (k, nl, dr, rate) in self.succ[l]:
                succ_i.append([k, nl, dr])
                succ_f.append(rate)
            succ_off[l + 1] = len(succ_i)
            for (c, ul, un, dl, dn, pos) in self.sl[l]:
                sl_i.append([c, ul, un, dl, dn, 0])
                sl_f.append(pos)
            sl_off[l + 1] = len(sl_i)
        ctrl_i, ctrl_f, seq_i, seq_f, sig_i, sig_f = [], [], [], [], [], []
        for c in self.ctrl:
            ctrl_i.append([len(seq_i), len(c["seqs"])])
            ctrl_f.append(c["cycle"])
            for s in c["seqs"]:
                seq_i.append([len(sig_i), len(s["signals"])])
                seq_f.append(s["len"])
                for (li, lo, g, a, d) in s["signals"]:
                    sig_i.append([li, lo])
                    sig_f.append([g, a, d])
        route_off = np.zeros(len(self.routes) + 1, np.int32)
        rl: List[int] = []
        for i, r in enumerate(self.routes):
            rl.extend(r)
            route_off[i + 1] = len(rl)
        entry_i, dem_f, dest_i, dest_f, dr_i, dr_f, lanecoef, typecoef = [], [], [], [], [], [], [], []
        for e in self.entries:
            nl = self.links[e["link"]][1]
            lc = e["lane_coefs"] or [1.0 / nl] * nl
            tc = e["type_coefs"] or ([1.0] + [0.0] * (T - 1))
            entry_i.append([e["node"], e["link"], len(dem_f), len(e["demand"]), len(dest_i), len(e["dests"]),
                            len(lanecoef), len(typecoef) | (int(e["exponential"]) << 30)])
            dem_f.extend([list(d) for d in e["demand"]])
            lanecoef.extend(lc)
            typecoef.extend(tc)
            for (ex, coeff, rts) in e["dests"]:
                dest_i.append([ex, len(dr_i), len(rts)])
                dest_f.append(coeff)
                for (r, rc) in rts:
                    dr_i.append(r)
                    dr_f.append(rc)

        def I(a, w):
            return np.asarray(a, np.int32).reshape(-1, w) if len(a) else np.zeros((0, w), np.int32)

        def D(a, w=1):
            return np.asarray(a, np.float64).reshape(-1, w) if len(a) else np.zeros((0, w), np.float64)

        def names(lst):
            return np.frombuffer(("\0".join(lst) + "\0").encode(), np.uint8)

        # merge points: convergents sorted by id like std::map<std::string, Convergent*> (reseau.cpp:14310)
        cv = sorted(self.cvgs, key=lambda c: c["name"].encode())
        cvg_i, cvg_f, cvg_tf, cvgam, pt_i, ptam_i = [], [], [], [], [], []
        for ci, c in enumerate(cv):
            cvg_i.append([c["node"], c["down_link"], len(pt_i), len(c["points"]), len(cvgam), len(c["am_links"]), 0, 0])
            cvg_f.append([c["tagreg"], c["gamma"], c["mu"], c["pos_cpt_av"]])
            cvg_tf.extend(c["tf"] or [0.0] * T)
            cvgam.extend(c["am_links"])
            for pt in c["points"]:
                pt_i.append([ci, pt["down_lane"], len(ptam_i), len(pt["vam"])])
                ptam_i.extend([[ln, niv] for (ln, niv) in pt["vam"]])
        mv_i, mv_l = [], []
        for (node, a, o, tl) in self.caf_mvts:
            mv_i.append([node, a, o, len(mv_l), len(tl)])
            mv_l.extend(tl)
        K = len(self.links)

        def csr(d):
            off = np.zeros(K + 1, np.int32)
            flat: List[int] = []
            for k in range(K):
                flat.extend(d.get(k, []))
                off[k + 1] = len(flat)
            return off, np.asarray(flat, np.int32)
        # distributors and plain merges: m_LstTuyAv of the downstream connection / m_LstTuyAm of the upstream one follow the allowed
        # movements (first appearance, lane order), like csrc/xmlnet.h; junctions and merges filled theirs when they were generated
        cam = {k: list(v) for k, v in self.link_cam.items()}
        cav = {k: list(v) for k, v in self.link_cav.items()}
        plain = {c["down_link"] for c in self.cvgs if c["node"] < 0}
        for k in range(K):
            if self.links[k][5] >= 0:
                continue
            first, nl = self.links[k][0], self.links[k][1]
            for ln in range(first, first + nl):
                for (key, nx, dr, rate) in self.succ[ln]:
                    o_ = self.lanes[nx][0]
                    if self.links[o_][5] >= 0:
                        continue
                    if o_ not in cav.setdefault(k, []):
                        cav[k].append(o_)
                    if o_ not in plain and k not in cam.setdefault(o_, []):
                        cam[o_].append(k)
        lcam_off, lcam = csr(cam)
        lcav_off, lcav = csr(cav)
        merge = {}
        if cv or len(lcam) or len(lcav):
            merge = {"cvg_i": I(cvg_i, 8), "cvg_f": D(cvg_f, 4), "cvg_tf": D(cvg_tf), "cvgam_i": np.asarray(cvgam, np.int32),
                     "pt_i": I(pt_i, 4), "ptam_i": I(ptam_i, 2), "cafmvt_i": I(mv_i, 5), "cafmvt_links": np.asarray(mv_l, np.int32),
                     "lcam_off": lcam_off, "lcam_links": lcam, "lcav_off": lcav_off, "lcav_links": lcav,
                     "names_cvg": names([c["name"] for c in cv])}

        return {**merge,
            "params": self.params, "cls": cls,
            "link_i": I(self.links, F.LINK_I), "link_f": D(self.link_f, F.LINK_F),
            "link_geom": D(self.link_geom, 4),
            "lane_i": I(self.lanes, F.LANE_I), "lane_f": D(self.lane_f),
            "succ_off": succ_off, "succ_i": I(succ_i, F.SUCC_I), "succ_f": D(succ_f),
            "sl_off": sl_off, "sl_i": I(sl_i, F.SL_I), "sl_f": D(sl_f),
            "ctrl_i": I(ctrl_i, 2), "ctrl_f": D(ctrl_f), "seq_i": I(seq_i, 2), "seq_f": D(seq_f),
            "sig_i": I(sig_i, 2), "sig_f": D(sig_f, 3),
            "node_i": I(self.nodes, F.NODE_I), "node_xy": D(self.node_xy, 2),
            "route_off": route_off, "route_links": np.asarray(rl, np.int32),
            "entry_i": I(entry_i, F.ENTRY_I), "dem_f": D(dem_f, 2), "dest_i": I(dest_i, F.DEST_I),
            "dest_f": D(dest_f), "droute_i": np.asarray(dr_i, np.int32), "droute_f": D(dr_f),
            "lanecoef_f": D(lanecoef), "typecoef_f": D(typecoef),
            "exit_i": I([[n, l] for (n, l, _) in self.exits], F.EXIT_I),
            "exit_f": D([c for (_, _, c) in self.exits]),
            "iv_i": I(self.iv_i, F.IV_I), "iv_f": D(self.iv_f, F.IV_F),
            "names_link": names(self.link_names), "names_node": names(self.node_names),
            "names_cls": names([c["name"] for c in self.classes]),
        }

    def write(self, path: str) -> None:
        F.write(path, self.sections())

    # -- the same network as a ROOT_SYMUBRUIT document (subset read by csrc/xmlnet.h) -------------
    def to_xml(self, path: str) -> None:
        """Reference-format input: what a SymuVia build (or `SymLoadNetworkEx` of this library) loads.  Junction-internal links, stop
        lines, merge points and their priority levels are NOT written: the loader derives them (CarrefourAFeuxEx::Init)."""
        from xml.sax.saxutils import quoteattr as q

        def num(x):
            return "infini" if x >= 1e9 else repr(float(x))

        def hms(sec):
            sec = float(sec)
            h, r = divmod(sec, 3600.0)
            m, s_ = divmod(r, 60.0)
            return f"{int(h):02d}:{int(m):02d}:{s_:09.6f}"
        T = len(self.classes)
        ext = [k for k in range(len(self.links)) if self.links[k][5] < 0]
        caf_nodes = sorted({self.links[k][5] for k in range(len(self.links)) if self.links[k][5] >= 0})
        plain_cvg = {c["name"]: c for c in self.cvgs if c["node"] < 0}
        o = ['<?xml version="1.0" encoding="UTF-8"?>', '<ROOT_SYMUBRUIT version="2.05">']
        P = self.params
        o.append(f' <SIMULATIONS><SIMULATION id="simu" pasdetemps={q(num(P[F.P_DT]))} debut="00:00:00" fin={q(hms(P[F.P_DT] * P[F.P_NSTEPS]))} loipoursuite="exacte" '
                 f'comportementflux="iti" seed="{int(P[F.P_SEED])}" date="2020-01-01"/></SIMULATIONS>')
        o.append(f' <TRAFICS><TRAFIC id="trafic" accbornee="{"true" if P[F.P_ACCBORNEE] else "false"}" coeffrelax={q(num(P[F.P_RELAX]))} chgtvoie="{"true" if P[F.P_CHGTVOIE] else "false"}" '
                 f'chgtvoie_dstfin={q(num(P[F.P_DSTCHGT]))} chgtvoie_dstfin_force={q(num(P[F.P_DST_FORCE]))} chgtvoie_dstfin_force_phi={q(num(P[F.P_PHI_FORCE] or 1.0))} traversees="false" '
                 'PeriodeAgregationCapteurs="30" Gamma="1" Mu="1" ti="0" pos_cpt_Av="5">')
        o.append('  <TYPES_DE_VEHICULE>')
        for c in self.classes:
            law = {F.LAW_NEWELL: "newell", F.LAW_GIPPS: "gipps", F.LAW_IDM: "idm"}[c["law"]]
            o.append(f'   <TYPE_DE_VEHICULE id={q(c["name"])} w={q(num(c["w"]))} kx={q(num(c["kx"]))} vx={q(num(c["vx"]))} espacement_arret={q(num(c["esp"]))} deceleration={q(num(c["decel"]))} '
                     f'chgtvoie_tau={q(num(c["tau_lc"]))} pi_rabattement={q(num(c["pi_rab"]))} loi="{law}"><ACCELERATION_PLAGES>'
                     + "".join(f'<ACCELERATION_PLAGE ax={q(num(ax))} vit_sup={q(num(vs))}/>' for (ax, vs) in c["plages"][:4]) + '</ACCELERATION_PLAGES></TYPE_DE_VEHICULE>')
        o.append('  </TYPES_DE_VEHICULE>')
        o.append('  <EXTREMITES>')
        exit_of_node = {n: (k, cap) for (n, k, cap) in self.exits}
        for e in self.entries:
            nl = self.links[e["link"]][1]
            lc = e["lane_coefs"] or [1.0 / nl] * nl
            tc = e["type_coefs"] or ([1.0] + [0.0] * (T - 1))
            kind = "distributionExponentielle" if e["exponential"] else "demande"
            o.append(f'   <EXTREMITE id={q(self.node_names[e["node"]])} typeCreationVehicule="{kind}"><FLUX_GLOBAL><FLUX><DEMANDES>'
                     + "".join(f'<DEMANDE niveau={q(num(lv))} duree={q(repr(float(du)))}/>' for (du, lv) in e["demand"]) + '</DEMANDES>'
                     f'<REP_VOIES><REP_VOIE coeffs={q(" ".join(repr(float(x)) for x in lc))}/></REP_VOIES>'
                     f'<REP_TYPEVEHICULES><REP_TYPEVEHICULE coeffs={q(" ".join(repr(float(x)) for x in tc))}/></REP_TYPEVEHICULES><REP_DESTINATIONS><REP_DESTINATION>'
                     + "".join(f'<DESTINATION sortie={q(self.node_names[self.exits[x][0]])} coeffOD={q(repr(float(co)))}>'
                               + "".join(f'<ROUTE id="R{r}" coeffAffectation={q(repr(float(rc)))}/>' for (r, rc) in rts) + '</DESTINATION>' for (x, co, rts) in e["dests"])
                     + '</REP_DESTINATION></REP_DESTINATIONS></FLUX></FLUX_GLOBAL></EXTREMITE>')
        for (n, k, cap) in self.exits:
            o.append(f'   <EXTREMITE id={q(self.node_names[n])}>' + (f'<CAPACITES><CAPACITE valeur={q(repr(float(cap)))}/></CAPACITES>' if cap >= 0 else '') + '</EXTREMITE>')
        o.append('  </EXTREMITES>')
        # allowed movements per connection
        o.append('  <CONNEXIONS_INTERNES>')
        by_node: Dict[int, Dict[int, list]] = {}
        for k in ext:
            first, nl, up, down = self.links[k][:4]
            for n_ in range(nl):
                ln = first + n_
                for (key, nx, dr, rate) in self.succ[ln]:
                    tgt = nx
                    if self.links[self.lanes[nx][0]][5] >= 0:          # junction-internal lane: the movement ends on the lane it leads to
                        tgt = self.succ[nx][0][1]
                    tl, tn = self.lanes[tgt][0], self.lanes[tgt][1]
                    ligne = 0.0
                    for (c_, ul, un, dl, dn, pos) in self.sl[ln]:
                        if dl == tl and dn == tn:
                            ligne = pos - self.lane_f[ln]
                    by_node.setdefault(down, {}).setdefault(ln, []).append((tl, tn, rate, ligne))
        for node in sorted(by_node):
            name = self.node_names[node]
            extra = ""
            cv = plain_cvg.get(name) or next((c for c in self.cvgs if c["node"] == node), None)
            if cv is not None:
                extra = f' PeriodeAgregationCapteurs={q(num(cv["tagreg"]))} Gamma={q(num(cv["gamma"]))} Mu={q(num(cv["mu"]))} pos_cpt_Av={q(num(cv["pos_cpt_av"]))}'
            o.append(f'   <CONNEXION_INTERNE id={q(name)}{extra}><MOUVEMENTS_AUTORISES>')
            for ln in sorted(by_node[node]):
                o.append(f'    <MOUVEMENT_AUTORISE id_troncon_amont={q(self.link_names[self.lanes[ln][0]])} num_voie_amont="{self.lanes[ln][1] + 1}"><MOUVEMENT_SORTIES>'
                         + "".join(f'<MOUVEMENT_SORTIE id_troncon_aval={q(self.link_names[tl])} num_voie_aval="{tn + 1}" taux_utilisation={q(repr(float(rate)))} ligne_feu={q(repr(float(lg)))}/>'
                                   for (tl, tn, rate, lg) in by_node[node][ln]) + '</MOUVEMENT_SORTIES></MOUVEMENT_AUTORISE>')
            o.append('   </MOUVEMENTS_AUTORISES>')
            if cv is not None and cv["tf"]:
                o.append('   <LISTE_TEMPS_INSERTION>' + "".join(f'<TEMPS_INSERTION id_typevehicule={q(self.classes[i]["name"])} ti={q(repr(float(t)))}/>' for i, t in enumerate(cv["tf"])) + '</LISTE_TEMPS_INSERTION>')
            o.append('   </CONNEXION_INTERNE>')
        o.append('  </CONNEXIONS_INTERNES>')
        o.append('  <CONTROLEURS_DE_FEUX>')
        for ci, c in enumerate(self.ctrl):
            o.append(f'   <CONTROLEUR_DE_FEUX id="CF{ci}" duree_vert_min="0" duree_rouge_degagement="0"><PLANS_DE_FEUX><PLAN_DE_FEUX id="P0" debut="00:00:00"><SEQUENCES>'
                     + "".join(f'<SEQUENCE duree_totale={q(repr(float(sq["len"])))}><SIGNAUX_ACTIFS>'
                               + "".join(f'<SIGNAL_ACTIF troncon_entree={q(self.link_names[li])} troncon_sortie={q(self.link_names[lo])} duree_vert={q(repr(float(g)))} duree_orange={q(repr(float(a)))} '
                                         f'duree_retard_allumage={q(repr(float(d)))}/>' for (li, lo, g, a, d) in sq["signals"]) + '</SIGNAUX_ACTIFS></SEQUENCE>' for sq in c["seqs"])
                     + '</SEQUENCES></PLAN_DE_FEUX></PLANS_DE_FEUX></CONTROLEUR_DE_FEUX>')
        o.append('  </CONTROLEURS_DE_FEUX>')
        o.append(' </TRAFIC></TRAFICS>')
        o.append(' <RESEAUX><RESEAU id="reseau">')
        o.append('  <TRONCONS>')
        for k in ext:
            first, nl, up, down = self.links[k][:4]
            g = self.link_geom[k]
            o.append(f'   <TRONCON id={q(self.link_names[k])} id_eltamont={q(self.node_names[up])} id_eltaval={q(self.node_names[down])} extremite_amont="{repr(float(g[0]))} {repr(float(g[1]))}" '
                     f'extremite_aval="{repr(float(g[2]))} {repr(float(g[3]))}" largeur_voie={q(repr(float(self.link_lane_w[k])))} nb_voie="{nl}" vit_reg={q(repr(float(self.link_f[k][1])))} '
                     f'longueur={q(repr(float(self.link_f[k][0])))}/>')
        o.append('  </TRONCONS>')

        def ctl(n):
            c = self.nodes[n][1]
            return f' controleur_de_feux="CF{c}"' if c >= 0 else ''
        o.append('  <CONNEXIONS>')
        o.append('   <EXTREMITES>' + "".join(f'<EXTREMITE id={q(self.node_names[n])}/>' for n in range(len(self.nodes)) if chr(self.nodes[n][0]) in "ES") + '</EXTREMITES>')
        o.append('   <REPARTITEURS>' + "".join(f'<REPARTITEUR id={q(self.node_names[n])}{ctl(n)}/>' for n in range(len(self.nodes)) if chr(self.nodes[n][0]) == "R") + '</REPARTITEURS>')
        o.append('   <CONVERGENTS>')
        for n in range(len(self.nodes)):
            if chr(self.nodes[n][0]) == "C" and n not in caf_nodes:
                cv = plain_cvg[self.node_names[n]]
                prio = {}
                for (ln, niv) in cv["points"][0]["vam"]:
                    prio[self.lanes[ln][0]] = niv
                o.append(f'    <CONVERGENT id={q(self.node_names[n])} id_TAv={q(self.link_names[cv["down_link"]])}><TRONCONS_AMONT>'
                         + "".join(f'<TRONCON_AMONT id={q(self.link_names[a])} priorite="{prio[a]}"/>' for a in cv["am_links"]) + '</TRONCONS_AMONT></CONVERGENT>')
        o.append('   </CONVERGENTS>')
        o.append('   <CARREFOURSAFEUX>')
        for n in caf_nodes:
            vmax = next(self.link_f[k][1] for k in range(len(self.links)) if self.links[k][5] == n)
            o.append(f'    <CARREFOURAFEUX id={q(self.node_names[n])} vit_max={q(repr(float(vmax)))}{ctl(n)}/>')
        o.append('   </CARREFOURSAFEUX>')
        o.append('  </CONNEXIONS>')
        o.append('  <ROUTES>')
        for r, links in enumerate(self.routes):
            o.append(f'   <ROUTE id="R{r}"><TRONCONS_>' + "".join(f'<TRONCON_ id={q(self.link_names[k])}/>' for k in links) + '</TRONCONS_></ROUTE>')
        o.append('  </ROUTES>')
        o.append('  <INIT><TRAJS>')
        for v, ((c, ln, rt, rp), (pos, vit, acc)) in enumerate(zip(self.iv_i, self.iv_f)):
            o.append(f'   <TRAJ id="{v}" tron={q(self.link_names[self.lanes[ln][0]])} voie="{self.lanes[ln][1] + 1}" dst={q(repr(float(pos)))} vit={q(repr(float(vit)))} acc={q(repr(float(acc)))} '
                     f'type_veh={q(self.classes[c]["name"])} route="R{rt}"/>')
        o.append('  </TRAJS></INIT>')
        o.append(' </RESEAU></RESEAUX>')
        o.append(' <SCENARIOS><SCENARIO id="scenario" simulation_id="simu" trafic_id="trafic" reseau_id="reseau" dirout="out" prefout="run"/></SCENARIOS>')
        o.append('</ROOT_SYMUBRUIT>')
        with open(path, "w", encoding="utf-8") as f:
            f.write("\n".join(o) + "\n")


# ---------------------------------------------------------------------------------------------
# C1: single link
# ---------------------------------------------------------------------------------------------
def single_link(length=2000.0, n_lanes=1, vit_reg=25.0, demand=0.5, n_steps=3600, dt=1.0, seed=42,
                law=F.LAW_NEWELL, exit_capacity=-1.0, signal=None, classes=None, type_coefs=None,
                exponential=False, accbornee=False) -> NetBuilder:
    """BASELINE config C1: Entree -> one link -> Sortie, deterministic demand (veh/s).

    `signal=(green, amber, red)` adds a mid-link pass-through repartiteur with a fixed-time
    stop line (two links of length/2), used by the signal known-answer tests.
    """
    b = NetBuilder(dt=dt, n_steps=n_steps, seed=seed, accbornee=accbornee)
    for c in (classes or [dict(VL, law=law)]):
        b.add_class(c)
    e = b.add_node("E", "E1", (0.0, 0.0))
    s = b.add_node("S", "S1", (length, 0.0))
    if signal is None:
        l1 = b.add_link("T1", e, s, n_lanes, length, vit_reg, "S", geom=(0, 0, length, 0))
        route = b.add_route([l1])
        last = l1
    else:
        g, a, r = signal
        ctrl = b.add_ctrl(g + a + r, [dict(len=g + a + r, signals=[])])
        rn = b.add_node("R", "R1", (length / 2, 0.0), ctrl=ctrl)
        l1 = b.add_link("T1", e, rn, n_lanes, length / 2, vit_reg, "R", geom=(0, 0, length / 2, 0))
        l2 = b.add_link("T2", rn, s, n_lanes, length / 2, vit_reg, "S", geom=(length / 2, 0, length, 0))
        b.ctrl[ctrl]["seqs"][0]["signals"].append((l1, l2, g, a, 0.0))
        for n in range(n_lanes):
            b.add_succ(b.lane_of(l1, n), l2, b.lane_of(l2, n), 1.0, 1)
            b.add_stopline(b.lane_of(l1, n), length / 2, ctrl, l1, n, l2, n)
        route = b.add_route([l1, l2])
        last = l2
    x = b.add_exit(s, last, exit_capacity)
    b.add_entry(e, l1, [(1e12, demand)], [(x, 1.0, [(route, 1.0)])], type_coefs=type_coefs,
                exponential=exponential)
    return b


# ---------------------------------------------------------------------------------------------
# C2 / C4 / C5: Manhattan grid of signalised junctions
# ---------------------------------------------------------------------------------------------
_DIRS = [(1, 0), (0, 1), (-1, 0), (0, -1)]       # E, N, W, S
_DNAME = "ENWS"


def grid(n=10, block=200.0, n_lanes=2, vit_reg=13.89, setback=12.0, lane_w=3.0, green=30.0, amber=3.0,
         n_steps=1800, dt=1.0, seed=42, demand_per_lane=0.2, demand_duration=1e12,
         preseed_per_lane=0.0, shuffle=False, classes=None, offsets=False, law=F.LAW_NEWELL,
         ny: Optional[int] = None, merges=True) -> NetBuilder:
    """n x ny grid of CAF junctions (BASELINE configs C2/C4/C5).

    Every approach lane k may go straight / right / left into exit lane k (no U-turn);
    signals: sequence 0 serves the E/W-bound approaches, sequence 1 the N/S-bound ones.
    `preseed_per_lane` > 0 places that many vehicles on every non-stub external lane with
    staircase routes to the boundary (C4/C5: no demand, population steady over the run).
    `merges`: every exit link fed by more than one internal link gets its `Convergent` (<CAF>_C0_<exit link>,
    CarrefourAFeuxEx.cpp:489-919) with one merge point per exit lane — row A10; False reproduces the round-1 tables.
    """
    ny = n if ny is None else ny
    rng = np.random.default_rng(seed)
    b = NetBuilder(dt=dt, n_steps=n_steps, seed=seed)
    for c in (classes or [dict(VL, law=law)]):
        b.add_class(c)
    cyc = 2 * (green + amber)
    J = {}
    for i in range(n):
        for j in range(ny):
            ctrl = b.add_ctrl(cyc, [dict(len=green + amber, signals=[]), dict(len=green + amber, signals=[])])
            J[(i, j)] = b.add_node("C", f"J{i}_{j}", (i * block, j * block), ctrl=ctrl)
    out_link: Dict[Tuple[int, int, int], int] = {}    # (i, j, d) link leaving junction (i,j) heading d
    in_link: Dict[Tuple[int, int, int], int] = {}     # (i, j, d) link arriving at junction (i,j) heading d
    link_dir: Dict[int, int] = {}
    stub_entries, stub_exits = [], []
    L = block - 2 * setback
    for i in range(n):
        for j in range(ny):
            x0, y0 = i * block, j * block
            for d, (dx, dy) in enumerate(_DIRS):
                i2, j2 = i + dx, j + dy
                gx0, gy0 = x0 + dx * setback, y0 + dy * setback
                if 0 <= i2 < n and 0 <= j2 < ny:
                    gx1, gy1 = i2 * block - dx * setback, j2 * block - dy * setback
                    k = b.add_link(f"L{i}_{j}{_DNAME[d]}", J[(i, j)], J[(i2, j2)], n_lanes, L, vit_reg, "D",
                                   geom=(gx0, gy0, gx1, gy1))
                    out_link[(i, j, d)] = k
                    in_link[(i2, j2, d)] = k
                    link_dir[k] = d
                else:
                    # boundary: exit stub leaving the grid and entry stub coming in
                    s = b.add_node("S", f"S{i}_{j}{_DNAME[d]}", (x0 + dx * block, y0 + dy * block))
                    k = b.add_link(f"X{i}_{j}{_DNAME[d]}", J[(i, j)], s, n_lanes, L, vit_reg, "S",
                                   geom=(gx0, gy0, x0 + dx * (block - setback), y0 + dy * (block - setback)))
                    out_link[(i, j, d)] = k
                    link_dir[k] = d
                    stub_exits.append((b.add_exit(s, k), k, i, j, d))
                    e = b.add_node("E", f"E{i}_{j}{_DNAME[d]}", (x0 + dx * block, y0 + dy * block))
                    din = (d + 2) % 4
                    k2 = b.add_link(f"N{i}_{j}{_DNAME[d]}", e, J[(i, j)], n_lanes, L, vit_reg, "D",
                                    geom=(x0 + dx * (block - setback), y0 + dy * (block - setback), gx0, gy0))
                    in_link[(i, j, din)] = k2
                    link_dir[k2] = din
                    stub_entries.append((e, k2, i, j, din))

    def lane_pt(link: int, num: int, end: bool) -> Tuple[float, float]:
        gx0, gy0, gx1, gy1 = b.link_geom[link]
        dx, dy = _DIRS[link_dir[link]]
        off = (num + 0.5) * lane_w
        px, py = (gx1, gy1) if end else (gx0, gy0)
        return px + dy * off, py - dx * off           # right-hand side of travel direction

    # CAF internal movements
    exit_in_count: Dict[int, int] = {}
    moves = []
    for (i, j), node in J.items():
        for din in range(4):
            a = in_link.get((i, j, din))
            if a is None:
                continue
            for turn in (0, 3, 1):                     # straight, right, left (relative)
                dout = (din + turn) % 4
                o = out_link.get((i, j, dout))
                if o is None:
                    continue
                moves.append((node, i, j, a, o, din))
                exit_in_count[o] = exit_in_count.get(o, 0) + n_lanes
    # Signal plan entries per (entry link, exit link); internal links in the creation order of CarrefourAFeuxEx::Init
    # (CarrefourAFeuxEx.cpp:331-480: m_mapMvtAutorises is ordered by entry-LANE label "<link>V<num+1>", then exit-link label, then
    # exit lane), which fixes m_LstMouvements, Divergent::m_LstTuyAv, Convergent::m_LstTuyAm and PointDeConvergence::m_LstVAm order.
    for (node, i, j, a, o, din) in moves:
        ctrl = b.nodes[node][1]
        seq = 0 if din in (0, 2) else 1
        b.ctrl[ctrl]["seqs"][seq]["signals"].append((a, o, green, amber, 0.0))
    by_node: Dict[int, list] = {}
    for (node, i, j, a, o, din) in moves:
        for k in range(n_lanes):
            by_node.setdefault(node, []).append((f"{b.link_names[a]}V{k + 1}", b.link_names[o], k, a, o, din))
    for node in sorted(by_node):
        ctrl = b.nodes[node][1]
        feeds: Dict[int, List[Tuple[int, int]]] = {}        # exit link -> [(internal link, exit lane num)] in creation order
        for (_, _, k, a, o, din) in sorted(by_node[node], key=lambda t: (t[0].encode(), t[1].encode(), t[2])):
            ax, ay = lane_pt(a, k, True)
            ox, oy = lane_pt(o, k, False)
            ln = math.sqrt((ox - ax) * (ox - ax) + (oy - ay) * (oy - ay))     # the same operations as csrc/xmlnet.h (no hypot: its rounding differs between libraries)
            ent_lane = b.lane_of(a, k)
            il = b.add_link(f"I{b.link_names[a]}_{b.link_names[o]}_{k + 1}", node, node, 1, ln, vit_reg,
                            "C" if exit_in_count[o] > 1 else "R", brick=node, prev_lane=ent_lane,
                            geom=(ax, ay, ox, oy))
            link_dir[il] = din
            b.add_succ(ent_lane, o, b.lane_of(il, 0), 1.0, 1)
            b.add_succ(b.lane_of(il, 0), o, b.lane_of(o, k), 1.0, 1)
            b.add_stopline(ent_lane, b.lane_f[ent_lane], ctrl, a, k, o, k)
            feeds.setdefault(o, []).append((il, k))
            b.caf_mvts.append((node, a, o, [il]))
            b.link_cam[il] = [a]                              # Divergent <CAF>_D0_<entry link>: one upstream link (:379-391)
            b.link_cav.setdefault(a, []).append(il)           # pDiv->AddEltAval(pT) (:707)
            b.link_cav[il] = [o]
        if not merges:
            continue
        for o, lst in feeds.items():
            b.link_cam[o] = [il for (il, _) in lst]           # pCvg->AddEltAmont(pT) (:708)
            if len(lst) < 2:
                continue                                      # one upstream link: converted to a Repartiteur (:860-905)
            points = []
            for kk in range(n_lanes):
                vam = [b.lane_of(il, 0) for (il, k) in lst if k == kk]
                # reference priority levels from the geometry (:1129-1166): upstream links sorted by AngleOfView seen from the
                # exit lane's upstream end (exit lane's downstream end as first point), smallest angle = level n (lowest)
                vx0, vy0 = lane_pt(o, kk, False)
                vx1, vy1 = lane_pt(o, kk, True)
                ang = {}
                for ln_ in vam:
                    gx0, gy0, _, _ = b.link_geom[b.lanes[ln_][0]]
                    a1, a2, b1, b2 = vx1 - vx0, vy1 - vy0, gx0 - vx0, gy0 - vy0
                    t = math.atan2(a1 * b2 - a2 * b1, a1 * b1 + a2 * b2) if (a1 * a1 + a2 * a2) and (b1 * b1 + b2 * b2) else 0.0
                    if t < 0:
                        t += 2 * math.pi
                    assert t not in ang, "two upstream links seen under the same angle: std::map insert would drop one"
                    ang[t] = ln_
                niv = {}
                lev = len(ang)
                for t in sorted(ang):
                    niv[ang[t]] = lev
                    lev -= 1
                points.append(dict(down_lane=b.lane_of(o, kk), vam=[(ln_, niv[ln_]) for ln_ in vam]))
            b.add_convergent(f"{b.node_names[node]}_C0_{b.link_names[o]}", node, o, [il for (il, _) in lst], points)

    # routing helpers on the (link -> link) graph without U-turns
    nxt: Dict[int, List[int]] = {}
    for (node, i, j, a, o, din) in moves:
        nxt.setdefault(a, []).append(o)

    if demand_per_lane > 0:
        # shortest paths entry stub -> exit stub (BFS over links; neighbour order straight, right, left)
        exit_of_link = {k: x for (x, k, _, _, _) in stub_exits}
        for (e, k0, i, j, din) in stub_entries:
            prev = {k0: -1}
            order = [k0]
            for cur in order:
                for o in nxt.get(cur, []):
                    if o not in prev:
                        prev[o] = cur
                        order.append(o)
            dests = []
            for (x, kx, xi, xj, xd) in stub_exits:
                if kx not in prev or (xi == i and xj == j):
                    continue
                path = [kx]
                while prev[path[-1]] != -1:
                    path.append(prev[path[-1]])
                dests.append((x, 1.0, [(b.add_route(path[::-1]), 1.0)]))
            tot = len(dests)
            dests = [(x, 1.0 / tot, r) for (x, _, r) in dests]
            b.add_entry(e, k0, [(demand_duration, demand_per_lane * n_lanes), (1e12, 0.0)], dests)

    if preseed_per_lane > 0:
        cls0 = b.classes[0]
        inv_kx = 1.0 / cls0["kx"]
        vmax = min(cls0["vx"], vit_reg)
        ext_links = [k for k in range(len(b.links)) if b.links[k][5] < 0 and chr(b.links[k][4]) == "D"
                     and b.link_names[k][0] == "L"]
        # two staircase routes per link: (d, d+1) and (d, d-1)
        link_at = {v: key for key, v in out_link.items()}
        route_ids: Dict[int, List[int]] = {}
        for k in ext_links:
            (i, j, d) = link_at[k]
            rts = []
            for side in (1, 3):
                d2 = (d + side) % 4
                path = [k]
                ci, cj = i + _DIRS[d][0], j + _DIRS[d][1]   # junction at the end of k
                cd = d
                while True:
                    # choose next heading among {d, d2} (monotone staircase => no repeated link)
                    cand = [dd for dd in (d, d2) if (ci, cj, dd) in out_link]
                    pick = cand[int(rng.integers(len(cand)))] if len(cand) > 1 else cand[0]
                    o = out_link[(ci, cj, pick)]
                    path.append(o)
                    if chr(b.links[o][4]) == "S":
                        break
                    ci, cj = ci + _DIRS[pick][0], cj + _DIRS[pick][1]
                    cd = pick
                rts.append(b.add_route(path))
            route_ids[k] = rts
        seeds = []
        for k in ext_links:
            for num in range(n_lanes):
                m = int(preseed_per_lane) + (1 if rng.random() < preseed_per_lane - int(preseed_per_lane) else 0)
                m = min(m, int(L / inv_kx) - 1)
                if m <= 0:
                    continue
                # uniform slots with jitter, min spacing 1/Kx
                slot = L / m
                jit = max(0.0, slot - inv_kx - 1e-6)
                pos = np.sort((np.arange(m) + 0.5) * slot - 0.5 * jit + rng.random(m) * jit * 0.999)
                for q in range(m):
                    s = (pos[q + 1] - pos[q]) if q + 1 < m else 1e9
                    v = min(vmax, max(0.0, -cls0["w"] * (s * cls0["kx"] - 1.0)))
                    seeds.append((k, num, float(pos[q]), float(v), route_ids[k][int(rng.integers(2))]))
        if shuffle:
            perm = rng.permutation(len(seeds))
            seeds = [seeds[p] for p in perm]
        for (k, num, p, v, r) in seeds:
            b.add_init_vehicle(0, b.lane_of(k, num), r, 0, p, v)
    return b


# ---------------------------------------------------------------------------------------------
# C3-style corridor: multi-lane, multi-class, lane changing, off-ramps reachable from the right lane only
# ---------------------------------------------------------------------------------------------
def corridor(n_links=10, length=500.0, n_lanes=3, vit_reg=25.0, demand=0.9, n_steps=600, seed=42, classes=None, type_coefs=None,
             offramp_every=0, offramp_share=0.1, preseed_density=0.0, chgtvoie=True, dst_chgtvoie=200.0, dst_force=0.0, phi_force=1.0,
             slow_lane_speed=None, shuffle=False, accbornee=False, relax=1.0, signal=None, signal_at=None, exponential=False,
             onramp_every=0, onramp_demand=0.15, onramp_length=250.0, onramp_tf=0.0, onramp_exponential=False) -> NetBuilder:
    """n_links consecutive links joined by pass-through distributors (lane k -> lane k).  Every `offramp_every` nodes a one-lane
    off-ramp leaves from lane 0 only, so exiting vehicles must reach lane 0 in the mandatory zone (`chgtvoie_dstfin`).
    Demand enters at the first link (lanes drawn uniformly); `preseed_density` (veh/m/lane) pre-seeds every lane.
    Every `onramp_every` nodes (offset by one from the off-ramps) the pass-through distributor is a `Convergent` instead (reseau.cpp:8300-8420):
    a one-lane on-ramp of priority level 2 joins the main road (level 1); every upstream lane is an upstream lane of every merge point
    (one per downstream lane), the ramp leads into lane 0.  The ramps have their own origins (`onramp_demand` veh/s each).
    `signal=(green, amber, red)` puts a fixed-time stop line on every lane at the end of link `signal_at` (default: the middle)."""
    rng = np.random.default_rng(seed)
    b = NetBuilder(n_steps=n_steps, seed=seed, chgtvoie=chgtvoie, dst_chgtvoie=dst_chgtvoie, accbornee=accbornee, relax=relax)
    b.params[F.P_PHI_FORCE] = phi_force
    b.params[F.P_DST_FORCE] = dst_force
    cl = classes or [VL, PL]
    for c in cl:
        b.add_class(c)
    tc = type_coefs or ([0.85, 0.15] if len(cl) == 2 else None)
    nodes = [b.add_node("E", "E0", (0.0, 0.0))]
    sig_j = (signal_at if signal_at is not None else n_links // 2) if signal else -1
    ctrl = b.add_ctrl(sum(signal), [dict(len=sum(signal), signals=[])]) if signal else -1
    on_at = [j for j in range(1, n_links) if onramp_every and j % onramp_every == (1 if onramp_every > 1 else 0) and j != sig_j]
    for j in range(1, n_links):
        nodes.append(b.add_node("C" if j in on_at else "R", f"{'CV' if j in on_at else 'R'}{j}", (j * length, 0.0), ctrl=ctrl if j == sig_j else -1))
    nodes.append(b.add_node("S", "S_end", (n_links * length, 0.0)))
    main = []
    for j in range(n_links):
        main.append(b.add_link(f"M{j}", nodes[j], nodes[j + 1], n_lanes, length, vit_reg, "S" if j == n_links - 1 else ("C" if j + 1 in on_at else "R"),
                               geom=(j * length, 0.0, (j + 1) * length, 0.0)))
    onramps = {}
    for j in on_at:
        e = b.add_node("E", f"EO{j}", (j * length - onramp_length, -60.0))
        r = b.add_link(f"O{j}", e, nodes[j], 1, onramp_length, vit_reg * 0.8, "C", geom=(j * length - onramp_length, -60.0, j * length, 0.0))
        b.add_succ(b.lane_of(r, 0), main[j], b.lane_of(main[j], 0), 1.0, 1)
        onramps[j] = (e, r)
        up = [b.lane_of(main[j - 1], k) for k in range(n_lanes)] + [b.lane_of(r, 0)]
        pts = [dict(down_lane=b.lane_of(main[j], l), vam=[(ln, 1) for ln in up[:-1]] + [(up[-1], 2)]) for l in range(n_lanes)]
        b.add_convergent(f"CV{j}", -1, main[j], [main[j - 1], r], pts, tf=[onramp_tf] * len(cl))
        b.link_cam[main[j]] = [main[j - 1], r]
        b.link_cav[main[j - 1]] = [main[j]]
        b.link_cav[r] = [main[j]]
    ramps = {}
    for j in range(1, n_links):
        for k in range(n_lanes):
            b.add_succ(b.lane_of(main[j - 1], k), main[j], b.lane_of(main[j], k), 1.0, 1)
        if j == sig_j:
            b.ctrl[ctrl]["seqs"][0]["signals"].append((main[j - 1], main[j], signal[0], signal[1], 0.0))
            for k in range(n_lanes):
                b.add_stopline(b.lane_of(main[j - 1], k), length, ctrl, main[j - 1], k, main[j], k)
        if offramp_every and j % offramp_every == 0:
            s = b.add_node("S", f"S{j}", (j * length, -100.0))
            r = b.add_link(f"X{j}", nodes[j], s, 1, 150.0, vit_reg * 0.6, "S", geom=(j * length, 0.0, j * length + 100.0, -100.0))
            b.add_succ(b.lane_of(main[j - 1], 0), r, b.lane_of(r, 0), 1.0, 1)
            ramps[j] = (r, b.add_exit(s, r))
    x_end = b.add_exit(nodes[-1], main[-1])
    through = b.add_route(main)
    dests = [(x_end, 1.0 - (offramp_share if ramps else 0.0), [(through, 1.0)])]
    ramp_routes = {}
    for j, (r, x) in ramps.items():
        ramp_routes[j] = b.add_route(main[:j] + [r])
        dests.append((x, offramp_share / len(ramps), [(ramp_routes[j], 1.0)]))
    if demand > 0:
        b.add_entry(nodes[0], main[0], [(1e12, demand)], dests, type_coefs=tc, exponential=exponential)
    for j, (e, r) in onramps.items():                                   # ramp origins: through traffic and the off-ramps further down
        dj = [(x_end, 1.0 - (offramp_share if any(jj > j for jj in ramps) else 0.0), [(b.add_route([r] + main[j:]), 1.0)])]
        later = [jj for jj in ramps if jj > j]
        for jj in later:
            dj.append((ramps[jj][1], offramp_share / len(later), [(b.add_route([r] + main[j:jj] + [ramps[jj][0]]), 1.0)]))
        if onramp_demand > 0:
            b.add_entry(e, r, [(1e12, onramp_demand)], dj, type_coefs=tc, exponential=onramp_exponential)
    if preseed_density > 0:
        seeds = []
        for j, k in enumerate(main):
            # routes starting at link j (through or a later ramp)
            later = [jj for jj in ramps if jj > j]
            for num in range(n_lanes):
                m = max(0, int(preseed_density * length))
                pos = np.sort(rng.random(m) * (length - 20.0) + 5.0)
                pos = pos[np.concatenate(([True], np.diff(pos) > 13.0))] if m else pos
                for q, p0 in enumerate(pos):
                    c = int(rng.random() < 0.15) if len(cl) > 1 else 0
                    if later and rng.random() < offramp_share:
                        jj = later[int(rng.integers(len(later)))]
                        rt = b.add_route(main[j:jj] + [ramps[jj][0]])
                    else:
                        rt = b.add_route(main[j:])
                    nxt = pos[q + 1] - p0 if q + 1 < len(pos) else 1e9
                    v = min(min(cl[c]["vx"], vit_reg), max(0.0, -cl[c]["w"] * (nxt * cl[c]["kx"] - 1.0)))
                    seeds.append((c, b.lane_of(k, num), rt, float(p0), float(v)))
        if shuffle:
            seeds = [seeds[p] for p in rng.permutation(len(seeds))]
        for (c, ln, rt, p0, v) in seeds:
            b.add_init_vehicle(c, ln, rt, 0, p0, v)
    return b


# ---------------------------------------------------------------------------------------------
# plain merge: two upstream links into one downstream link through a Convergent (reseau.cpp:6432-6473, :8300-8420)
# ---------------------------------------------------------------------------------------------
def ymerge(len_main=400.0, len_ramp=300.0, len_down=500.0, lanes_main=1, lanes_down=1, demand_main=0.3, demand_ramp=0.2, tf=0.0, vit_reg=20.0,
           vit_ramp=None, n_steps=400, exit_capacity=-1.0, classes=None, type_coefs=None, prio_main=1, prio_ramp=2, seed=42, exponential=False,
           tagreg=30.0, gamma=1.0, mu=1.0, pos_cpt_av=5.0, accbornee=False) -> NetBuilder:
    """Main road (priority level `prio_main`) and a one-lane ramp (`prio_ramp`) merging into a downstream link; main lane k -> downstream lane
    min(k, lanes_down - 1), ramp -> lane 0.  Every upstream lane is an upstream lane of every merge point, as the loader builds them."""
    b = NetBuilder(n_steps=n_steps, seed=seed, accbornee=accbornee)
    cl = classes or [VL]
    for c in cl:
        b.add_class(c)
    em = b.add_node("E", "E_main", (0.0, 0.0))
    er = b.add_node("E", "E_ramp", (len_main - len_ramp, -80.0))
    m = b.add_node("C", "CV", (len_main, 0.0))
    x = b.add_node("S", "S_end", (len_main + len_down, 0.0))
    lm = b.add_link("MAIN", em, m, lanes_main, len_main, vit_reg, "C", geom=(0.0, 0.0, len_main, 0.0))
    lr = b.add_link("RAMP", er, m, 1, len_ramp, vit_ramp or vit_reg, "C", geom=(len_main - len_ramp, -80.0, len_main, 0.0))
    ld = b.add_link("DOWN", m, x, lanes_down, len_down, vit_reg, "S", geom=(len_main, 0.0, len_main + len_down, 0.0))
    for k in range(lanes_main):
        b.add_succ(b.lane_of(lm, k), ld, b.lane_of(ld, min(k, lanes_down - 1)), 1.0, 1)
    b.add_succ(b.lane_of(lr, 0), ld, b.lane_of(ld, 0), 1.0, 1)
    up = [(b.lane_of(lm, k), prio_main) for k in range(lanes_main)] + [(b.lane_of(lr, 0), prio_ramp)]
    b.add_convergent("CV", -1, ld, [lm, lr], [dict(down_lane=b.lane_of(ld, l), vam=list(up)) for l in range(lanes_down)],
                     tagreg=tagreg, gamma=gamma, mu=mu, pos_cpt_av=pos_cpt_av, tf=[tf] * len(cl))
    b.link_cam[ld] = [lm, lr]
    b.link_cav[lm] = [ld]
    b.link_cav[lr] = [ld]
    ex = b.add_exit(x, ld, exit_capacity)
    if demand_main > 0:
        b.add_entry(em, lm, [(1e12, demand_main)], [(ex, 1.0, [(b.add_route([lm, ld]), 1.0)])], type_coefs=type_coefs, exponential=exponential)
    if demand_ramp > 0:
        b.add_entry(er, lr, [(1e12, demand_ramp)], [(ex, 1.0, [(b.add_route([lr, ld]), 1.0)])], type_coefs=type_coefs, exponential=exponential)
    return b


def config(name: str, **kw) -> NetBuilder:
    """Named BASELINE.json configurations."""
    if name == "C1":
        return single_link(**kw)
    if name == "C2":
        return grid(n=10, n_steps=1800, demand_per_lane=0.2, demand_duration=1200.0, **kw)
    if name == "C3":    # 10 km arterial, 3 lanes, 2 classes, lane changing, on-ramp Convergents and off-ramp Repartiteurs every 1 km, ~0.06 veh/m/lane pre-seeded
        kw.setdefault("n_steps", 600)
        km = kw.pop("km", 10)      # BASELINE.json quotes ~100 k vehicles: 0.06 veh/m/lane x 3 lanes needs ~550 km of such an arterial (km=550); km=10 is the geometry it names
        return corridor(n_links=2 * km, length=500.0, n_lanes=3, demand=1.2, offramp_every=2, offramp_share=0.1, onramp_every=2, onramp_demand=0.2,
                        preseed_density=0.06, chgtvoie=True, **kw)
    if name == "C4":
        return grid(n=100, n_steps=300, demand_per_lane=0.0, preseed_per_lane=12.6, **kw)
    if name == "C5":
        return grid(n=300, n_steps=100, demand_per_lane=0.0, preseed_per_lane=14.0, **kw)
    raise KeyError(name)


def fast_grid(path: str, nx: int, ny: int, per_lane: float = 12.6, seed: int = 42, route_len: int = 24, n_steps: int = 300,
              rank: int = 0, world: int = 1, merges: bool = True):
    """C++ generator (csrc/gen_grid.cpp) of the pre-seeded grid: the whole network, the vehicles of strip `rank` of `world`
    (global ids).  Returns (vehicles written, vehicles in the whole population)."""
    import ctypes as C
    import os
    lib = C.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "libsymgen.so"))
    lib.symgen_grid2.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_longlong), C.c_int]
    tot = C.c_longlong(0)
    n = lib.symgen_grid2(path.encode(), nx, ny, per_lane, seed, route_len, n_steps, rank, world, C.byref(tot), int(merges))
    if n < 0:
        raise RuntimeError(f"symgen_grid could not write {path}")
    return n, int(tot.value)
