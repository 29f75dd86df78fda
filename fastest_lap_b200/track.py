"""Track XML ("discrete" format) -> per-node constants for the collocation kernels.

Host-side mirror of `create_track_from_xml` (src/main/c/fastestlapc.cpp:176-203):
  Circuit_preprocessor(Xml_document&) loader      src/core/applications/circuit_preprocessor.hpp:64-258
  Track_by_polynomial(const Circuit_preprocessor&) src/core/vehicles/track_by_polynomial.hpp:78-153
The reference builds order-1 (piecewise linear) polynomials of the arclength for the centreline, yaw, yaw',
pitch, roll, pitch', roll', nl and nr; closed tracks get one closing sample (L, value[0]) appended, the yaw included
(no 2 pi unwrap, track_by_polynomial.hpp:115).  pitch'/roll' are fed from the XML arrays dpitch_dot/droll_dot
(track_by_polynomial.hpp:87-88), which is what the reference's golden files were produced with.
The track is evaluated in plain doubles and never differentiated (road_curvilinear.hpp:104-135): everything the
kernels need is a per-node constant, computed here once per mesh.
"""
import xml.etree.ElementTree as ET
import numpy as np


def _vec(node):
    return np.array([float(t) for t in node.text.replace("\n", " ").split(",") if t.strip()], dtype=np.float64)


class Track:
    def __init__(self, s, x, y, z, yaw, pitch, roll, yaw_dot, dpitch, droll, nl, nr, track_length, closed, with_elevation, name=""):
        self.name = name
        self.closed = bool(closed)
        self.with_elevation = bool(with_elevation)
        self.track_length = float(track_length)
        self.s_data = np.asarray(s, dtype=np.float64)
        arrays = dict(x=x, y=y, z=z, yaw=yaw, pitch=pitch, roll=roll, yaw_dot=yaw_dot, dpitch=dpitch, droll=droll, nl=nl, nr=nr)
        sk = self.s_data
        if self.closed:
            sk = np.append(sk, self.track_length)
        self._s = sk
        self._a = {}
        for k, v in arrays.items():
            v = np.asarray(v, dtype=np.float64)
            if self.closed:
                v = np.append(v, v[0])
            self._a[k] = v

    @property
    def n_points(self):
        return len(self.s_data)

    def has_elevation(self):
        # track_by_polynomial.hpp:156-160
        a = self._a
        return bool(np.any(a["pitch"] != 0.0) or np.any(a["roll"] != 0.0) or np.any(np.abs(a["z"]) >= 1.0e-12)
                    or np.any(a["dpitch"] != 0.0) or np.any(a["droll"] != 0.0))

    def __call__(self, key, s):
        """Piecewise-linear evaluation (lion sPolynomial of order 1)."""
        return np.interp(np.asarray(s, dtype=np.float64), self._s, self._a[key])

    def left_track_limit(self, s):
        return self("nl", s)

    def right_track_limit(self, s):
        return self("nr", s)

    def node_data(self, s):
        """[n][6] array: theta, mu, phi, dtheta/ds, dmu/ds, dphi/ds at the mesh nodes `s`."""
        s = np.asarray(s, dtype=np.float64)
        out = np.empty((len(s), 6))
        for c, k in enumerate(("yaw", "pitch", "roll", "yaw_dot", "dpitch", "droll")):
            out[:, c] = self(k, s)
        return out

    def position(self, s):
        return np.stack([self("x", s), self("y", s), self("z", s)], axis=-1)


def track_from_xml(source, name="", use_kerbs=True):
    text = source if source.lstrip().startswith("<") else open(source).read()
    root = ET.fromstring(text)
    if root.attrib.get("format") != "discrete":
        raise ValueError('Track should be of type "discrete"')
    ttype = root.attrib.get("type")
    if ttype not in ("open", "closed"):
        raise ValueError('Incorrect track type, should be "open" or "closed"')
    dims = root.attrib.get("dimensions")
    if dims not in ("2", "3"):
        raise ValueError("Track XML file must have a dimensions attribute equal to 2 or 3")
    elev = dims == "3"
    data = root.find("data")
    n = int(data.attrib["number_of_points"])
    s = _vec(data.find("arclength"))
    x = _vec(data.find("centerline/x"))
    y = _vec(data.find("centerline/y"))
    z = _vec(data.find("centerline/z")) if elev else np.zeros(n)
    yaw = _vec(data.find("yaw"))
    yaw_dot = _vec(data.find("yaw_dot"))
    nl = _vec(data.find("nl"))
    nr = _vec(data.find("nr"))
    if elev:
        pitch, roll = _vec(data.find("pitch")), _vec(data.find("roll"))
        dpitch, droll = _vec(data.find("dpitch_dot")), _vec(data.find("droll_dot"))
    else:
        pitch = roll = dpitch = droll = np.zeros(n)
    for a in (s, x, y, z, yaw, yaw_dot, nl, nr, pitch, roll, dpitch, droll):
        if len(a) != n:
            raise ValueError("track arrays must have number_of_points entries")
    length = float(root.find("header/track_length").text)
    # kerbs (track_by_polynomial.hpp:93-108; XML layout kerb.h:158-170: every child of <left> / <right> is a kerb with
    # arclength_start, arclength_finish, width and a `used` attribute): a sample inside a used kerb gets nl / nr widened by
    # the width of the FIRST such kerb
    kerbs = root.find("kerbs") if use_kerbs else None
    if kerbs is not None:
        for side, arr in (("left", nl), ("right", nr)):
            node = kerbs.find(side)
            if node is None:
                continue
            done = np.zeros(len(s), dtype=bool)
            for kerb in node:
                if kerb.attrib.get("used", "").strip().lower() not in ("1", "true"):
                    continue
                s0, s1, wd = float(kerb.findtext("arclength_start")), float(kerb.findtext("arclength_finish")), float(kerb.findtext("width"))
                mask = (s > s0 - 1.0e-12) & (s < s1 + 1.0e-12) & ~done
                arr[mask] += wd
                done |= mask
    return Track(s, x, y, z, yaw, pitch, roll, yaw_dot, dpitch, droll, nl, nr, length, ttype == "closed", elev, name)


def track_to_xml(track):
    """The track as a database file of the reference's discrete format (the one track_from_xml and the C entry
    create_track_from_xml read: track_by_polynomial.hpp:78-153)."""
    n = track.n_points
    fmt = lambda a: ", ".join(repr(float(v)) for v in np.asarray(a)[:n])
    elev = track.with_elevation
    root = ET.Element("circuit", format="discrete", type="closed" if track.closed else "open", dimensions="3" if elev else "2")
    ET.SubElement(ET.SubElement(root, "header"), "track_length").text = repr(float(track.track_length))
    data = ET.SubElement(root, "data", number_of_points=str(n))
    ET.SubElement(data, "arclength").text = fmt(track.s_data)
    cl = ET.SubElement(data, "centerline")
    ET.SubElement(cl, "x").text = fmt(track._a["x"]); ET.SubElement(cl, "y").text = fmt(track._a["y"])
    if elev:
        ET.SubElement(cl, "z").text = fmt(track._a["z"])
    for tag, key in (("yaw", "yaw"), ("yaw_dot", "yaw_dot"), ("nl", "nl"), ("nr", "nr")) + ((("pitch", "pitch"), ("roll", "roll"), ("dpitch_dot", "dpitch"), ("droll_dot", "droll")) if elev else ()):
        ET.SubElement(data, tag).text = fmt(track._a[key])
    return ET.tostring(root, encoding="unicode")


_NPZ_KEYS = ("x", "y", "z", "yaw", "pitch", "roll", "yaw_dot", "dpitch", "droll", "nl", "nr")


def track_to_npz(track, path):
    """Compact binary form of a loaded track (the arrays of the XML, before the closing sample is appended)."""
    n = track.n_points
    arrays = {k: track._a[k][:n] for k in _NPZ_KEYS}
    np.savez_compressed(path, s=track.s_data, track_length=track.track_length, closed=int(track.closed),
                        with_elevation=int(track.with_elevation), **arrays)


def track_from_npz(path, name=""):
    d = np.load(path)
    return Track(d["s"], *[d[k] for k in _NPZ_KEYS], float(d["track_length"]), bool(int(d["closed"])), bool(int(d["with_elevation"])), name)


class TrackByArcs:
    """Straight-and-corner tracks of the reference's tests (src/core/vehicles/track_by_arcs.hpp:9-184; XML
    `<track type="straight-and-curve" width=...>` with `<straight length=.../>` and `<corner sweep radius direction/>`
    children): analytic centreline, constant half width `scale * width / 2` on both sides, no elevation.  Left-hand
    corners have NEGATIVE curvature in the file's convention (track_by_arcs.hpp:42-43); the heading rate handed to the
    road is `dot((-r'_y, r'_x), r'') / |r'|^2` (:127-137), which is what `node_data` returns.  On closed tracks the last
    segment is not read but closes the loop (:69-76).  Same interface as `Track`."""

    def __init__(self, source, scale=1.0, closed=True, name=""):
        text = source if source.lstrip().startswith("<") else open(source).read()
        root = ET.fromstring(text)
        segs = list(root)
        if not segs:
            raise ValueError("Track is empty, and it cannot be used.")
        self.name, self.closed = name, bool(closed)
        self.half_width = 0.5 * scale * float(root.attrib["width"])
        n = len(segs)
        P, D, S0, K = np.zeros((n, 2)), np.tile([1.0, 0.0], (n, 1)), np.zeros(n), np.zeros(n)
        total = 0.0

        def corner(el):
            R, theta = scale * float(el.attrib["radius"]), float(el.attrib["sweep"])
            return (-1.0 / R if el.attrib["direction"].strip() == "left" else 1.0 / R), R, theta

        for i, el in enumerate(segs[:-1]):
            if el.tag == "straight":
                L = scale * float(el.attrib["length"])
                P[i + 1], D[i + 1], S0[i + 1] = P[i] + D[i] * L, D[i], S0[i] + L
                total += L
            elif el.tag == "corner":
                k, R, theta = corner(el)
                K[i] = k
                nrm = np.array([-D[i][1], D[i][0]])
                c2s = -nrm / k
                centre = P[i] - c2s
                d1, d2 = c2s / np.linalg.norm(c2s), D[i] / np.linalg.norm(D[i])
                P[i + 1] = centre + R * (d1 * np.cos(theta) + d2 * np.sin(theta))
                S0[i + 1] = S0[i] + R * theta
                D[i + 1] = d2 * np.cos(theta) - d1 * np.sin(theta)
                total += R * theta
            else:
                raise ValueError("Node not recognized")
        last = segs[-1]
        if last.tag == "straight":
            total += np.linalg.norm(P[-1] - P[0]) if closed else scale * float(last.attrib["length"])
        elif last.tag == "corner":
            k, R, theta = corner(last)
            K[-1] = k
            total += R * theta
        else:
            raise ValueError("Node not recognized")
        self._P, self._D, self._S0, self._K = P, D, S0, K
        self.track_length = float(total)

    def has_elevation(self):
        return False

    def _at(self, s):
        """r, r', r'' at arclength s (track_by_arcs.hpp:140-184)."""
        i = 0
        while i < len(self._S0) - 1 and s >= self._S0[i + 1]:
            i += 1
        s0, k, p, d = s - self._S0[i], self._K[i], self._P[i], self._D[i]
        if abs(k) < 1.0e-10:
            return p + s0 * d, d.copy(), np.zeros(2)
        theta = s0 * abs(k)
        nrm = np.array([-d[1], d[0]]) / k
        centre = p + nrm
        d1 = -nrm / np.linalg.norm(nrm)
        dr = d * np.cos(theta) - d1 * np.sin(theta)
        return centre + (d1 * np.cos(theta) + d * np.sin(theta)) / abs(k), dr, np.array([-dr[1], dr[0]]) * k

    def __call__(self, key, s):
        s = np.atleast_1d(np.asarray(s, dtype=np.float64))
        out = np.zeros(len(s))
        for j, sj in enumerate(s):
            r, dr, d2r = self._at(float(sj))
            if key == "x":
                out[j] = r[0]
            elif key == "y":
                out[j] = r[1]
            elif key == "yaw":
                out[j] = np.arctan2(dr[1], dr[0])
            elif key == "yaw_dot":
                out[j] = (-dr[1] * d2r[0] + dr[0] * d2r[1]) / (dr[0] * dr[0] + dr[1] * dr[1])
            elif key in ("nl", "nr"):
                out[j] = self.half_width
            elif key not in ("z", "pitch", "roll", "dpitch", "droll"):
                raise KeyError(key)
        return out

    def left_track_limit(self, s):
        return self("nl", s)

    def right_track_limit(self, s):
        return self("nr", s)

    node_data = Track.node_data
    position = Track.position
