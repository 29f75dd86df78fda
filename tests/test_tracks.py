"""Straight-and-corner tracks (track_by_arcs.hpp of the reference, used by its oval-track tests)."""
import numpy as np
import pytest

from fastest_lap_b200.track import TrackByArcs
from tests.helpers import load_golden

OVAL = """<ovaltrack type="straight-and-curve" width="20.0">
    <straight length="150.0"/>
    <corner sweep="1.570796326794897" radius="50.0" direction="left"/>
    <straight length="100.0"/>
    <corner sweep="1.570796326794897" radius="90.0" direction="left"/>
    <straight length="300.0"/>
    <corner sweep="1.570796326794897" radius="50.0" direction="left"/>
    <straight length="100.0"/>
    <corner sweep="1.570796326794897" radius="90.0" direction="left"/>
    <straight length="150.0"/>
</ovaltrack>"""


def test_oval_closes_and_matches_the_golden_mesh():
    """The oval of the reference's tests (src/test/applications/data/ovaltrack.xml, restated above): four quarter corners
    sum to a full turn, the last straight closes the loop, and the node data of the golden fixtures (made from the
    reference's own file) is reproduced."""
    for name, scale in (("f1_ovaltrack_closed", 1.0), ("kart_ovaltrack_derivative", 0.2)):
        p, ex = load_golden(name)
        t = TrackByArcs(OVAL, scale, True)
        assert abs(t.track_length - p.track_length) <= 1e-9 * p.track_length
        assert abs(t.track_length - scale * (800.0 + 0.5 * np.pi * 280.0)) <= 1e-9 * t.track_length
        np.testing.assert_allclose(t.node_data(p.s), p.track_nodes, rtol=0, atol=1e-12)
        np.testing.assert_allclose(t.left_track_limit(p.s), p.wl, atol=0)
        assert np.all(p.wl == 10.0 * scale) and np.all(p.wr == 10.0 * scale)
        end = t.position(np.array([t.track_length * (1.0 - 1e-12)]))[0]
        assert np.linalg.norm(end) <= 1e-6
    t = TrackByArcs(OVAL, 1.0, True)
    s = np.linspace(0.0, t.track_length, 2001)[:-1]
    kappa = t("yaw_dot", s)
    assert set(np.round(kappa, 12)) == {0.0, -0.02, round(-1.0 / 90.0, 12)}          # left-hand corners: negative curvature
    assert abs(np.trapezoid(kappa, s) + 2.0 * np.pi) <= 5e-2          # piecewise-constant curvature sampled every 0.6 m
    assert not t.has_elevation()


def test_malformed_tracks_are_refused():
    with pytest.raises(ValueError):
        TrackByArcs('<t type="straight-and-curve" width="1.0"></t>')
    with pytest.raises(ValueError):
        TrackByArcs('<t type="straight-and-curve" width="1.0"><straight length="1.0"/><chicane/><straight length="1.0"/></t>')


def test_kerbs_widen_the_track_limits():
    """track_by_polynomial.hpp:93-108 with the XML layout of kerb.h:158-170: used kerbs widen nl / nr on their stretch (the
    first matching kerb only), unused ones do nothing."""
    from fastest_lap_b200.track import track_from_xml
    n = 11
    arr = lambda v: ", ".join(repr(float(x)) for x in v)
    s = np.linspace(0.0, 100.0, n)
    zeros = arr(np.zeros(n))
    xml = ("<circuit format=\"discrete\" type=\"open\" dimensions=\"2\"><header><track_length>100.0</track_length></header>"
           "<data number_of_points=\"%d\"><arclength>%s</arclength><centerline><x>%s</x><y>%s</y></centerline><yaw>%s</yaw><yaw_dot>%s</yaw_dot><nl>%s</nl><nr>%s</nr></data>"
           "<kerbs><left><kerb_element_1 used=\"1\"><arclength_start>20.0</arclength_start><arclength_finish>40.0</arclength_finish><width>1.5</width></kerb_element_1>"
           "<kerb_element_2 used=\"1\"><arclength_start>30.0</arclength_start><arclength_finish>60.0</arclength_finish><width>0.5</width></kerb_element_2></left>"
           "<right><kerb_element_1 used=\"0\"><arclength_start>0.0</arclength_start><arclength_finish>100.0</arclength_finish><width>2.0</width></kerb_element_1></right></kerbs></circuit>"
           % (n, arr(s), arr(s), zeros, zeros, zeros, arr(np.full(n, 5.0)), arr(np.full(n, 6.0))))
    t = track_from_xml(xml)
    np.testing.assert_allclose(t.left_track_limit(s), [5, 5, 6.5, 6.5, 6.5, 5.5, 5.5, 5, 5, 5, 5])
    np.testing.assert_allclose(t.right_track_limit(s), 6.0)
