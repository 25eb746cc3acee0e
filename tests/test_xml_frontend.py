"""ROOT_SYMUBRUIT front end (csrc/xmlnet.h behind SymLoadNetworkEx / symgpu_create): CPU only.

The generators write every network twice: as flat tables (the device's input) and as a reference-format XML document
(`NetBuilder.to_xml`) in which the junction-internal links, stop lines, merge points and their priority
levels are NOT spelled out.  The XML front end must rebuild them like the reference's loader (CarrefourAFeuxEx::Init,
Convergent::Init): the flattened tables it produces are compared with the generator's own, and the oracle run on both files
must give the same trajectories, draw for draw."""
import ctypes as C
import importlib

import numpy as np
import pytest

from oracle.binding import Oracle

F = importlib.import_module("open-symuvia_b200.flat")


def _cases(synth):
    return {
        "c1": (synth.single_link(n_steps=300), 300),
        "c1_signal_capacity": (synth.single_link(n_steps=300, signal=(20, 3, 17), vit_reg=15.0, demand=0.4, exit_capacity=0.3), 300),
        "grid3": (synth.grid(n=3, n_steps=300, demand_per_lane=0.3, demand_duration=200, preseed_per_lane=5, shuffle=True), 300),
        "grid_1lane_2classes": (synth.grid(n=3, ny=2, n_lanes=1, n_steps=200, demand_per_lane=0.3, classes=[synth.VL, synth.PL]), 200),
        "ymerge_multilane": (synth.ymerge(lanes_main=2, lanes_down=2, demand_main=0.8, tf=2.0, classes=[synth.VL, synth.PL], type_coefs=[0.8, 0.2]), 300),
        "corridor_ramps": (synth.corridor(n_links=6, length=400.0, n_lanes=3, demand=1.0, offramp_every=2, onramp_every=2, preseed_density=0.04,
                                          shuffle=True, n_steps=200), 200),
    }


@pytest.mark.parametrize("name", ["c1", "c1_signal_capacity", "grid3", "grid_1lane_2classes", "ymerge_multilane", "corridor_ramps"])
def test_xml_document_flattens_to_the_generators_tables(pkg, synth, tmp_path, name):
    b, steps = _cases(synth)[name]
    pa, px, pb = str(tmp_path / "a.symflat"), str(tmp_path / "a.xml"), str(tmp_path / "b.symflat")
    b.write(pa)
    b.to_xml(px)
    L = pkg.lib()
    L.symflat_from_xml.argtypes = [C.c_char_p, C.c_char_p]
    assert L.symflat_from_xml(px.encode(), pb.encode()) == 0
    A, B = F.read(pa), F.read(pb)
    # node numbering differs (the document groups connections by kind); everything indexed by link, lane, route or class must be equal
    for sec in ("cls", "link_f", "link_geom", "lane_i", "lane_f", "succ_off", "succ_i", "succ_f", "sl_off", "sl_f", "ctrl_f", "seq_i", "seq_f", "sig_i", "sig_f",
                "route_off", "route_links", "dem_f", "dest_f", "droute_i", "droute_f", "lanecoef_f", "typecoef_f", "iv_i", "iv_f", "cvg_f", "cvg_tf", "cvgam_i", "pt_i",
                "ptam_i", "cafmvt_links", "lcam_off", "lcam_links", "lcav_off", "lcav_links"):
        if sec in A or sec in B:
            assert sec in A and sec in B and A[sec].shape == B[sec].shape and np.array_equal(A[sec], B[sec]), sec
    assert np.array_equal(A["link_i"].reshape(-1, 8)[:, [0, 1, 4]], B["link_i"].reshape(-1, 8)[:, [0, 1, 4]])      # lanes, lane count, downstream connection type
    oa, ob = Oracle(pa, len(b.links), len(b.classes)), Oracle(pb, len(b.links), len(b.classes))
    for t in range(steps):
        assert oa.step() == ob.step()
        sa, sb = oa.state(), ob.state()
        for key in ("id", "lane", "route_pos", "flags", "leader", "pos", "vit", "acc", "deltaN"):
            assert np.array_equal(sa[key], sb[key]), (t, key)
        assert oa.rng_count == ob.rng_count
    assert oa.n > 0


def test_xml_errors_are_reported(pkg, tmp_path):
    L = pkg.lib()
    L.symflat_from_xml.argtypes = [C.c_char_p, C.c_char_p]
    bad = tmp_path / "bad.xml"
    bad.write_text("<ROOT_SYMUBRUIT><SIMULATIONS/></ROOT_SYMUBRUIT>")
    assert L.symflat_from_xml(str(bad).encode(), str(tmp_path / "o.symflat").encode()) != 0
    assert L.symflat_from_xml(str(tmp_path / "missing.xml").encode(), str(tmp_path / "o.symflat").encode()) != 0


def test_unsupported_features_are_refused(pkg, synth, tmp_path):
    """Roundabouts and conflict points inside junctions change the hot path: a document that uses them must not load as if it did not."""
    L = pkg.lib()
    L.symflat_from_xml.argtypes = [C.c_char_p, C.c_char_p]
    px = tmp_path / "a.xml"
    synth.grid(n=2, n_steps=50, demand_per_lane=0.2).to_xml(str(px))
    doc = px.read_text()
    out = str(tmp_path / "o.symflat").encode()
    assert L.symflat_from_xml(str(px).encode(), out) == 0
    assert "<CONNEXIONS>" in doc
    g = tmp_path / "g.xml"
    g.write_text(doc.replace("<CONNEXIONS>", "<CONNEXIONS><GIRATOIRES><GIRATOIRE id=\"G1\"/></GIRATOIRES>", 1))
    assert L.symflat_from_xml(str(g).encode(), out) != 0
    if 'traversees="false"' in doc:
        t = tmp_path / "t.xml"
        t.write_text(doc.replace('traversees="false"', 'traversees="true"', 1))
        assert L.symflat_from_xml(str(t).encode(), out) != 0
