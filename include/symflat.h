/* symflat.h — SYMFLAT1: flattened network / demand tables.
 *
 * The reference builds a pointer graph at load time (Reseau::LoadReseauXML,
 * symuvia/src/reseau.cpp:4931-9612; CarrefourAFeuxEx::Init, CarrefourAFeuxEx.cpp:215-1196).
 * The B200 path needs the same information as dense tables.  This container is what
 * `SymLoadNetworkEx("*.symflat")` reads directly and what the XML front end produces in memory.
 *
 * File: "SYMFLAT1" | u32 n_sections | { char name[24] | u32 dtype | u64 count | data (pad to 8 B) }*
 * dtype 0 = int32, 1 = float64, 2 = uint8.  All ids are dense 0-based indices.
 *
 * section      dtype rows x stride   meaning (reference origin)
 * params       f64   32              SYMFLAT_P_* below (SIMULATION/TRAFIC attributes, reseau.cpp:5330-5625)
 * cls          f64   T x 24          vehicle types: w kx vx espacement_arret deceleration chgtvoie_tau pi_rabattement
 *                                    law n_plages (ax,vit_sup)x4        (reseau.cpp:5640-5900)
 * link_i       i32   K x 8           first_lane n_lanes up_node down_node type_aval brick(-1 | CAF node) 0 0
 * link_f       f64   K x 2           length vit_reg                      (tuyau.h)
 * lane_i       i32   L x 4           link num prev_lane flags(bit0 = mandatory lane change, voie.cpp:789)
 *                                    prev_lane: entry lane of the movement for a CAF internal lane, else -1
 * lane_f       f64   L               length
 * succ_off     i32   L + 1           CSR of allowed movements per lane   (Connection.h m_mapMvtAutorises)
 * succ_i       i32   M x 3           key_link(next external link of the itinerary) next_lane flags
 * succ_f       f64   M               taux_utilisation
 * sl_off       i32   L + 1           CSR of stop lines per lane          (voie.h:60-80 LigneDeFeu)
 * sl_i         i32   S x 6           ctrl up_link up_lane_num down_link down_lane_num 0
 * sl_f         f64   S               position on the lane
 * ctrl_i/f     i32 C x 2 / f64 C     seq_off n_seq / cycle length        (ControleurDeFeux.h PlanDeFeux)
 * seq_i/f      i32 Q x 2 / f64 Q     sig_off n_sig / total length        (CDFSequence)
 * sig_i/f      i32 G x 2 / f64 G x 3 link_in link_out / green amber delay (SignalActif)
 * node_i       i32 Nn x 4            type('E','S','R','C') ctrl 0 0
 * route_off    i32   R + 1           CSR of itineraries (external links only)
 * route_links  i32
 * entry_i      i32   E x 8           node link dem_off n_dem dest_off n_dest lanecoef_off typecoef_off|(exponential<<30)
 * dem_f        f64   * x 2           duration level(veh/s)               (DEMANDES/DEMANDE)
 * dest_i/f     i32 * x 3 / f64       exit droute_off n_routes / coeffOD
 * droute_i/f   i32 / f64             route id / coeffAffectation
 * lanecoef_f   f64                   REP_VOIE coefficients ; typecoef_f: type repartition
 * exit_i/f     i32 X x 2 / f64 X     node link / capacity veh/s (<0 = unlimited, sortie.cpp:60-118)
 * iv_i/f       i32 N0 x 4 / f64 x 3  class lane route route_pos / pos vit acc  (INIT/TRAJS, reseau.cpp:11680-11773)
 * merge points (row A10; all optional — a network without Convergent objects has none):
 * cvg_i        i32   V x 8           node(CAF brick | -1) down_link pt_off n_pts am_off n_am 0 0   (Convergent, convergent.cpp:159-255;
 *                                    rows in Reseau::Liste_convergents order = std::map by id, reseau.h:201)
 * cvg_f        f64   V x 4           Tagreg gamma mu pos_cpt_aval ; cvg_tf f64 V x T: insertion time per vehicle type (GetTf)
 * cvgam_i      i32                   upstream links (m_LstTuyAm order): one sensor each at length - 5
 * pt_i         i32   P x 4           cvg down_lane vam_off n_vam                     (PointDeConvergence: one per downstream lane)
 * ptam_i       i32   * x 2           upstream lane, reference priority level (m_LstVAm order; 1 = highest)
 * cafmvt_i     i32   * x 5           node entry_link exit_link tl_off n_tl ; cafmvt_links: internal links (MvtCAFEx::lstTMvt)
 * lcam_off/links  i32 K+1 / *        upstream links of a link's upstream connection (m_LstTuyAm), lcav_*: downstream links of
 *                                    its downstream connection (m_LstTuyAv), both in the reference's order
 * names_link / names_node / names_cls / names_cvg   u8 '\0'-separated labels (XML output)
 * link_geom    f64   K x 4           x0 y0 x1 y1 (abs/ord of <TRAJ>)
 */
#ifndef SYMFLAT_H
#define SYMFLAT_H

#define SYMFLAT_MAGIC "SYMFLAT1"
enum { SYMFLAT_P_DT = 0, SYMFLAT_P_NSTEPS, SYMFLAT_P_SEED, SYMFLAT_P_ACCBORNEE, SYMFLAT_P_RELAX,
       SYMFLAT_P_CHGTVOIE, SYMFLAT_P_DSTCHGT, SYMFLAT_P_TRAVERSEES, SYMFLAT_P_GHOST, SYMFLAT_P_PHI_FORCE,
       SYMFLAT_P_DST_FORCE, SYMFLAT_P_COUNT = 32 };
enum { SYMFLAT_CLS_STRIDE = 24, SYMFLAT_LAW_NEWELL = 0, SYMFLAT_LAW_GIPPS = 1, SYMFLAT_LAW_IDM = 2 };

#endif
