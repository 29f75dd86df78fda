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
