"""The SymuVia C exports (SymLoadNetworkEx / SymRunNextStepEx / ... ) driven the way a ctypes user of the
reference drives them, checked against the oracle at the 2-decimal precision of the <TRAJ> output
(XMLDocTrafic.cpp:2232-2257).  Needs a B200."""
import re

import numpy as np
import pytest

from oracle.binding import Oracle

pytestmark = pytest.mark.gpu

TRAJ = re.compile(r'<TRAJ id="(\d+)" tron="([^"]+)" voie="(\d+)" dst="([-\d.]+)" abs="[-\d.]+" ord="[-\d.]+" z="0.00" '
                  r'vit="([-\d.]+)" acc="([-\d.]+)" type="(\w+)"/>')


def _expected(o, b):
    s = o.state()
    rows = []
    for i in range(len(s["id"])):
        if s["pos"][i] < 0:
            continue
        ln = int(s["lane"][i])
        rows.append((str(int(s["id"][i])), b.link_names[b.lanes[ln][0]], str(b.lanes[ln][1] + 1), "%.2f" % s["pos"][i],
                     "%.2f" % s["vit"][i], "%.2f" % s["acc"][i], b.classes[int(s["flags"][i]) >> 8]["name"]))
    return rows


def test_step_exports_match_oracle(pkg, synth, tmp_path):
    b = synth.grid(n=3, n_steps=80, demand_per_lane=0.3, preseed_per_lane=4, shuffle=True)
    p = str(tmp_path / "g.symflat")
    b.write(p)
    o = Oracle(p, len(b.links))
    sv = pkg.SymuVia()
    assert sv.step()[0] is False                      # no network loaded (Symubruit.cpp:146-150)
    assert sv.load(str(tmp_path / "missing.symflat")) is False
    assert sv.load(p) is True
    for k in range(80):
        ok, end, xml = sv.step()
        n = o.step()
        assert ok and end == (k == 79)
        assert f'nbVeh="{n}"' in xml and xml.startswith("<INST val=") and xml.endswith("</INST>")
        assert TRAJ.findall(xml) == _expected(o, b), f"step {k + 1}"
    assert sv.step()[0] is False                      # networks are released at the end of the run (:1010-1011)


def test_lite_and_run_exports(pkg, synth, tmp_path):
    b = synth.single_link(length=500.0, n_steps=50)
    p = str(tmp_path / "c.symflat")
    b.write(p)
    sv = pkg.SymuVia()
    assert sv.load(p)
    n = 0
    while True:
        ok, end, xml = sv.step(lite=True)
        assert ok and xml is None
        n += 1
        if end:
            break
    assert n == 50
    ok, end, _ = sv.step(lite=True)                   # the Lite overload leaves the network loaded at the end of the run (Symubruit.cpp:1051-1060)
    assert ok and end
    sv.unload()
    assert sv.step(lite=True)[0] is False
    assert sv.run(p) == 0 and sv.run(str(tmp_path / "nope.symflat")) == 1


def test_load_reference_format_xml(pkg, synth, tmp_path):
    """SymLoadNetworkEx on a ROOT_SYMUBRUIT document (csrc/xmlnet.h rebuilds the junction-internal links, stop lines and merge points):
    same <TRAJ> output as the oracle on the generator's own flat tables, on a signalised grid with merges and on a ramp merge."""
    for b, steps in ((synth.grid(n=3, n_steps=100, demand_per_lane=0.3, preseed_per_lane=4, shuffle=True), 100),
                     (synth.ymerge(lanes_main=2, lanes_down=2, demand_main=0.8, n_steps=120, classes=[synth.VL, synth.PL], type_coefs=[0.8, 0.2]), 120)):
        p, x = str(tmp_path / "g.symflat"), str(tmp_path / "g.xml")
        b.write(p)
        b.to_xml(x)
        o = Oracle(p, len(b.links), len(b.classes))
        sv = pkg.SymuVia()
        assert sv.load(x) is True
        for k in range(steps):
            ok, end, xml = sv.step()
            n = o.step()
            assert ok and end == (k == steps - 1)
            assert f'nbVeh="{n}"' in xml
            assert TRAJ.findall(xml) == _expected(o, b), f"step {k + 1}"


def test_create_vehicle_export(pkg, synth, tmp_path):
    """SymCreateVehicleEx (Symubruit.cpp:4345 -> reseau.cpp:13351): ids, error codes and trajectories."""
    b = synth.single_link(length=900.0, vit_reg=20.0, demand=0.2, n_steps=120, classes=[synth.VL, synth.PL], type_coefs=[1.0, 0.0])
    p = str(tmp_path / "c.symflat")
    b.write(p)
    sv = pkg.SymuVia()
    assert sv.create_vehicle("VL", "E1", "S1", 1, 0.5) == -1          # no network
    assert sv.load(p)
    o = Oracle(p, len(b.links), 2)
    assert sv.create_vehicle("XX", "E1", "S1", 1, 0.5) == -2
    assert sv.create_vehicle("VL", "E9", "S1", 1, 0.5) == -3
    assert sv.create_vehicle("VL", "E1", "S9", 1, 0.5) == -4
    assert sv.create_vehicle("VL", "E1", "S1", 3, 0.5) == -5
    for k in range(120):
        if k in (3, 4, 5, 40, 41):                                     # bursts: the second/third must queue at the origin
            cls = 1 if k in (4, 40) else 0
            ida = sv.create_vehicle(b.classes[cls]["name"], "E1", "S1", 1, 0.25)
            idb = o.create_vehicle(cls, 0, 0, 0, 0.25)
            assert ida == idb >= 0
        ok, end, xml = sv.step()
        o.step()
        assert ok
        assert TRAJ.findall(xml) == _expected(o, b), f"step {k + 1}"
    sv.unload()
