"""Decode the event log written by oracle/ref_stage_harness.cpp into id-level narrowphase facts.

The harness records, in call order, every hot-path predicate call the unmodified reference made during
one mcDispatch (kind codes below).  Coordinates are resolved to polygon-soup ids through the ps arrays
the harness dumped from the reference's own `ps` half-edge mesh.

kinds: 1 plane_coefficients | 2 segment_plane_type (start) | 7 segment_plane_type (result)
       3 orient3d | 4 orient2d | 5 segment_plane_intersection | 6 point_in_polygon (3D)
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np


def split_events(ev: np.ndarray) -> List[Tuple[int, np.ndarray]]:
    out = []
    i = 0
    n = ev.size
    while i < n:
        kind = int(ev[i])
        ln = int(ev[i + 1])
        out.append((kind, ev[i + 2:i + 2 + ln]))
        i += 2 + ln
    return out


def _key3(p) -> bytes:
    return np.asarray(p, dtype=np.float64).tobytes()


class SoupIndex:
    """coordinates -> edge id / face id, from the reference's ps dump.  Keys are coordinate bit patterns (some corpus meshes
    hold several vertices at one position, so a position does not name a vertex; an edge or a face is still named by the
    positions of its vertices unless two of them coincide entirely: naming such a one raises)."""

    def __init__(self, ps_xyz, ps_face_sizes, ps_face_vtx, ps_edges):
        self.xyz = ps_xyz
        key = [_key3(p) for p in ps_xyz]
        self.edge: Dict[Tuple[bytes, bytes], int] = {}
        for e, (vs, vt, _f0, _f1) in enumerate(ps_edges):
            k = (key[int(vs)], key[int(vt)])
            self.edge[k] = e if k not in self.edge else -1  # ambiguous: an error only if a test names it
        self.face: Dict[Tuple[bytes, ...], int] = {}
        off = 0
        for f, n in enumerate(ps_face_sizes):
            k = tuple(key[int(v)] for v in ps_face_vtx[off:off + n])
            self.face[k] = f if k not in self.face else -1
            off += int(n)

    def e(self, q, r) -> int:
        e = self.edge[(_key3(q), _key3(r))]  # the reference always passes source(h0), target(h0)
        if e < 0:
            raise ValueError("two polygon-soup edges share both end positions; cannot resolve ids from coordinates")
        return e

    def f(self, verts) -> int:
        f = self.face[tuple(_key3(p) for p in verts)]
        if f < 0:
            raise ValueError("two polygon-soup faces share all vertex positions; cannot resolve ids from coordinates")
        return f


def decode_dispatch(events: np.ndarray, idx: SoupIndex):
    """Returns (planes, tests): planes = {face: (normal[3], d, max_comp)};
    tests = list of dicts {edge, face, type, sign_q, sign_r, o3d (list of results), pip (list of chars),
    point (or None)} in the reference's call order."""
    planes: Dict[int, Tuple[np.ndarray, float, int]] = {}
    tests: List[dict] = []
    cur = None
    in_type = False
    for kind, d in split_events(events):
        if kind == 1:
            n = int(d[0])
            verts = d[1:1 + 3 * n].reshape(n, 3)
            normal = d[1 + 3 * n:4 + 3 * n].copy()
            dd = float(d[4 + 3 * n])
            mc = int(d[5 + 3 * n])
            planes[idx.f(verts)] = (normal, dd, mc)
        elif kind == 2:
            q, r = d[0:3], d[3:6]
            n = int(d[6])
            verts = d[7:7 + 3 * n].reshape(n, 3)
            cur = {"edge": idx.e(q, r), "face": idx.f(verts), "type": None, "o3d": [], "o2d": [], "pip": [],
                   "point": None, "nverts": n}
            tests.append(cur)
            in_type = True
        elif kind == 7:
            cur["type"] = chr(int(d[0]))
            in_type = False
        elif kind == 3:
            if in_type:
                cur["o3d"].append(float(d[12]))
        elif kind == 4:
            if in_type:
                cur["o2d"].append(float(d[6]))
        elif kind == 5:
            if cur is not None and cur["type"] == "1" and cur["point"] is None:
                cur["point"] = d[10:13].copy()
        elif kind == 6:
            if cur is None:
                continue
            # grammar (kernel.cpp:2515-2575): '1' -> exactly one PIP after the plane point; 'q'/'r' -> one;
            # 'p' -> up to two (stops at the first decisive one); anything later belongs to downstream stages
            t = cur["type"]
            limit = {"1": 1, "q": 1, "r": 1, "p": 2, "0": 0}.get(t, 0)
            if t == "1" and cur["point"] is None:
                continue
            if len(cur["pip"]) < limit:
                if t == "p" and cur["pip"] and cur["pip"][-1] in "ive":
                    continue
                n = int(d[3])
                verts = d[4:4 + 3 * n].reshape(n, 3)
                if idx.f(verts) == cur["face"]:
                    cur["pip"].append(chr(int(d[8 + 3 * n])))
    for t in tests:
        s = [int(np.sign(x)) for x in t["o3d"][-2:]] if len(t["o3d"]) >= 2 else [0, 0]
        t["sign_q"], t["sign_r"] = s[0], s[1]
    return planes, tests
