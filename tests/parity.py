"""Shared parity criteria (BASELINE.json north_star): hit primitive index and inside flag exact, hit distance and
normal within 1e-5 relative (f64 mode) / 1e-4 (f32 mode).

In f32 mode a ray whose two nearest candidates are closer together than the tolerance (coplanar faces, shared
edges) is legitimately ambiguous: such rays may report the other primitive, provided the reported distance agrees
with the oracle's within the tolerance. They are counted and bounded separately."""
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# measured counts of the last check_hits call (ambiguous / unresolvable rays of the f32 comparison), and a log of all of
# them for the run (gpurun_out/parity_stats.jsonl comes back from the GPU box): the caps below are set from these
LAST = {}


def _log(label, stats):
    LAST.clear()
    LAST.update(stats)
    print("parity[%s]: %s" % (label, json.dumps(stats)))
    try:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "parity_stats.jsonl"), "a") as f:
            f.write(json.dumps(dict(label=label, **stats)) + "\n")
    except OSError:
        pass


def silhouette_distance(arrays, prim, o, d):
    """How close (relative to the primitive's size) the ray o + t d passes to the outline of primitive `prim`: the smallest
    barycentric margin for a triangle / parallelogram, |distance to the centre - r| / r for a sphere. A ray within f32
    resolution of an outline may legitimately land on the other side of it in f32 mode."""
    kind, g = arrays["kind"][prim], arrays["geom"][prim]
    if kind == 0:  # RTC_KIND_TRIANGLE
        v0, e1, e2 = g[0:3], g[3:6], g[6:9]
        off = o - v0
        s1, s2 = np.cross(off, e1), np.cross(d, e2)
        det = np.dot(e1, s2)
        if det == 0:
            return 0.0
        u, v = np.dot(off, s2) / det, np.dot(d, s1) / det
        mirror = bool(arrays["flags"][prim] & 1)
        return float(min(abs(u), abs(v), abs(1 - u), abs(1 - v)) if mirror else min(abs(u), abs(v), abs(1 - u - v)))
    if kind == 1 and arrays["xform"][prim] < 0:  # plain sphere
        off = o - g[0:3]
        dn = d / np.linalg.norm(d)
        perp = off - np.dot(off, dn) * dn
        return float(abs(np.linalg.norm(perp) - g[3]) / g[3])
    return 1.0


def check_hits(got, want, tol, exact, max_ambiguous_frac=0.0005, normal_tol=None, origins=None, dirs=None,
               max_unresolvable_frac=0.0, label="", skip=None, arrays=None, max_silhouette_frac=0.0):
    """exact=True (f64 mode): primitive and inside flag bit-exact, tolerance relative to t itself. exact=False (f32 mode): the
    ray origin is only known to 2^-24 relative, so t cannot be better than that times the origin's magnitude; the distance
    tolerance is therefore relative to max(|t|, |origin|) when `origins` is given. Returns the number of ambiguous rays.
    RTC_PARITY_MEASURE=1 turns the count caps into a log (gpurun_out/parity_stats.jsonl) -- how the caps were set."""
    measure = bool(os.environ.get("RTC_PARITY_MEASURE"))
    normal_tol = tol if normal_tol is None else normal_tol
    n = len(want)
    assert len(got) == n
    same = (got["prim"] == want["prim"]) & (got["inside"] == want["inside"])
    hit = want["prim"] >= 0
    scale = np.maximum(np.abs(want["t"]), 1e-300)
    if not exact and origins is not None:
        scale = np.maximum(scale, np.linalg.norm(origins, axis=1))
    t_ok = np.abs(got["t"] - want["t"]) <= tol * scale
    n_unres = ambiguous = hard = n_sil = 0
    if exact:
        assert same.all(), "primitive/inside mismatch on %d of %d rays (first: %s vs %s)" % (
            (~same).sum(), n, got[~same][:1], want[~same][:1])
    else:
        bad = ~same
        if origins is not None:
            # Two classes of rays are below the resolution of an f32 origin and are excluded from the f32 comparison (counted
            # and bounded; they only arise for rays that start on a surface):
            #  (a) grazing re-hits of the skip hit's own primitive. The reference's self-hit rule accepts a second hit on the
            #      same primitive as soon as it lies 1e-12 (relative) away from the origin (Util.NearEnough = 1e-24 on squared
            #      distances); for a sphere the chord of such a re-hit is only resolved to ~sqrt(2^-24) of the radius in f32.
            #      Criterion: oracle hit on the skip primitive within 4 x tol of the origin (`skip` given);
            #  (b) hits closer to the origin than the f32 origin itself is known (16 ulp of its magnitude): the rounded origin
            #      may lie on either side of that surface.
            omag = np.maximum(np.linalg.norm(origins, axis=1), 1.0)
            unresolvable = hit & (np.abs(want["t"]) <= 16 * 2.0 ** -24 * omag)
            if skip is not None:
                unresolvable |= hit & (skip["prim"] >= 0) & (want["prim"] == skip["prim"]) & (np.abs(want["t"]) <= 4 * tol * omag)
            n_unres = int(unresolvable.sum())
            bad &= ~unresolvable
            same = same | unresolvable
            got = got.copy()
            got[unresolvable] = want[unresolvable]
            t_ok = t_ok | unresolvable
        # an ambiguous ray must still be a hit at the same distance
        legit = bad & hit & (got["prim"] >= 0) & t_ok
        ambiguous = int(legit.sum())
        #  (c) silhouette rays: the ray passes the outline of the oracle's (or the device's) primitive closer than f32 resolves
        #      (relative margin < 1e-4): one side sees the primitive, the other looks past it. Needs the scene's arrays.
        if arrays is not None and origins is not None and dirs is not None:
            for i in np.nonzero(bad & ~legit)[0]:
                margin = min(silhouette_distance(arrays, int(p), origins[i], dirs[i]) for p in (want["prim"][i], got["prim"][i]) if p >= 0)
                if margin < 1e-4:
                    legit[i] = True
                    n_sil += 1
        hard = int((bad & ~legit).sum())
    m = hit & same
    t_bad = int((~t_ok[m]).sum())
    dn = np.linalg.norm(got["normal"][m] - want["normal"][m], axis=1)
    ntol = np.full(dn.shape, normal_tol)
    if not exact and dirs is not None:
        # f32 mode: on curved primitives a hit point moves along the ray by dt ~ eps_f32 * scale / cos(incidence), and the
        # normal with it; at grazing incidence (|cos| < 0.1) rounding the sphere's own centre/radius to f32 already moves
        # the hit by ~sqrt(eps) of the radius, so those hits get the looser bound 20 * tol.
        cosi = np.abs(np.sum(dirs[m] * want["normal"][m], axis=1)) / np.maximum(np.linalg.norm(dirs[m], axis=1), 1e-300)
        ntol = np.where(cosi >= 0.1, normal_tol, 20 * normal_tol)
    n_bad = int((dn > ntol).sum())
    pos_scale = np.maximum(np.linalg.norm(want["position"][m], axis=1), 1.0)
    dp = np.linalg.norm(got["position"][m] - want["position"][m], axis=1)
    p_bad = int((dp > 10 * tol * pos_scale).sum())
    stats = dict(n=int(n), hits=int(hit.sum()), exact=bool(exact), ambiguous=ambiguous, unresolvable=n_unres, silhouette=n_sil, mismatch=hard,
                 t_off=t_bad, normal_off=n_bad, position_off=p_bad,
                 cap_ambiguous=float(max_ambiguous_frac * n + 2), cap_unresolvable=float(max_unresolvable_frac * n + 2))
    def ex(idx):
        return [dict(i=int(i), got=[int(got["prim"][i]), int(got["inside"][i]), float(got["t"][i])] + [float(v) for v in got["normal"][i]],
                     want=[int(want["prim"][i]), int(want["inside"][i]), float(want["t"][i])] + [float(v) for v in want["normal"][i]],
                     o=[float(v) for v in origins[i]] if origins is not None else None,
                     d=[float(v) for v in dirs[i]] if dirs is not None else None) for i in idx[:3]]
    if hard and not exact:
        stats["first_mismatch"] = ex(np.nonzero(bad & ~legit)[0])
    mi = np.nonzero(m)[0]
    if t_bad:
        stats["first_t_off"] = ex(mi[~t_ok[m]])
    if n_bad:
        stats["first_normal_off"] = ex(mi[dn > ntol])
    _log(label, stats)
    if measure:
        return ambiguous
    assert n_unres <= max_unresolvable_frac * n + 2, "too many sub-resolution hits: %d of %d" % (n_unres, n)
    assert n_sil <= max_silhouette_frac * n + 2, "too many silhouette rays: %d of %d" % (n_sil, n)
    assert hard == 0, "non-ambiguous primitive mismatch on %d rays: %s" % (hard, stats.get("first_mismatch"))
    assert ambiguous <= max_ambiguous_frac * n + 2, "too many ambiguous rays: %d of %d" % (ambiguous, n)
    assert t_bad == 0, "hit distance off by more than %g relative on %d rays: max %g" % (
        tol, t_bad, np.max(np.abs(got["t"][m] - want["t"][m]) / scale[m]))
    assert n_bad == 0, "normal off by more than %g on %d rays: max %g" % (normal_tol, n_bad, (dn / ntol).max() * normal_tol)
    assert p_bad == 0, "hit position off on %d rays: max %g" % (p_bad, dp.max())
    return ambiguous


def random_rays(rng, n, lo, hi, ray_dt):
    rays = np.zeros(n, ray_dt)
    rays["origin"] = rng.uniform(lo, hi, (n, 3))
    d = rng.normal(size=(n, 3))
    rays["dir"] = d / np.linalg.norm(d, axis=1, keepdims=True)
    return rays
