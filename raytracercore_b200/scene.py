"""Scene — Python view of the C++ host layer's Scene / SceneLoader (raytracercore_b200/host/), which mirror the
reference's Raytracing/Scene.cs and SceneLoader.cs. Pure host code: works without a GPU."""
import ctypes as C

import numpy as np

from . import _native as N


class LoaderException(RuntimeError):
    """SceneLoader.cs:16-26 — message carries the command and 1-based line number."""


class Scene:
    def __init__(self, handle):
        if not handle:
            raise ValueError("null scene handle")
        self._h = C.c_void_p(handle)

    # -- construction (SceneLoader.FromFile, SceneLoader.cs:112) -------------------------------------------
    @staticmethod
    def from_file(path):
        err = C.create_string_buffer(1024)
        h = N.lib.rtcs_scene_load(str(path).encode(), err, len(err))
        if not h:
            if err.value:
                raise LoaderException(err.value.decode())
            return None  # the reference returns null for a missing file (SceneLoader.cs:430-439)
        return Scene(h)

    @staticmethod
    def from_string(text):
        err = C.create_string_buffer(1024)
        h = N.lib.rtcs_scene_parse(text.encode(), err, len(err))
        if not h:
            raise LoaderException(err.value.decode())
        return Scene(h)

    @staticmethod
    def synthetic(name, n, seed, jitter=0.0):
        """BASELINE.json synthetic scenes: 'soup' (n triangles) / 'spheres' (n spheres)."""
        h = N.lib.rtcs_scene_synthetic(name.encode(), int(n), int(seed), float(jitter))
        if not h:
            raise ValueError("unknown synthetic scene %r" % name)
        return Scene(h)

    def close(self):
        if self._h:
            N.lib.rtcs_scene_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- Scene fields (Scene.cs:16-35) -------------------------------------------------------------------
    def globals(self):
        g = N.Globals()
        rc = N.lib.rtcs_scene_globals(self._h, C.byref(g))
        assert rc == 0
        return g

    @property
    def width(self):
        return self.globals().width

    @property
    def height(self):
        return self.globals().height

    @property
    def recursion(self):
        return self.globals().recursion

    @property
    def n_prims(self):
        return self.globals().n_prims

    @property
    def n_cameras(self):
        return self.globals().n_cameras

    def override(self, width=-1, height=-1, recursion=-1, camera=-1):
        rc = N.lib.rtcs_scene_override(self._h, width, height, recursion, camera)
        if rc:
            raise ValueError("invalid scene override")

    def set_ambient(self, rgb):
        a = (C.c_double * 3)(*rgb)
        N.lib.rtcs_scene_set_ambient(self._h, a)

    def set_debug_geom(self, on):
        N.lib.rtcs_scene_set_debug_geom(self._h, 1 if on else 0)

    # -- flattened views ---------------------------------------------------------------------------------
    def desc(self):
        d = N.SceneDesc()
        rc = N.lib.rtcs_scene_desc(self._h, C.byref(d))
        assert rc == 0
        return d

    def arrays(self):
        """Copies of the flattened primitive arrays as numpy (kind, flags, geom[n,12], xform, xforms[m,48], material[n,14])."""
        d = self.desc()
        n, m = d.n_prims, d.n_xforms

        def arr(ptr, count, dtype):
            if count == 0 or not ptr:
                return np.zeros(0, dtype=dtype)
            return np.ctypeslib.as_array(ptr, shape=(count,)).astype(dtype, copy=True)

        return dict(kind=arr(d.kind, n, np.uint8), flags=arr(d.flags, n, np.uint8),
                    geom=arr(d.geom, n * 12, np.float64).reshape(n, 12),
                    xform=arr(d.xform, n, np.int32), xforms=arr(d.xforms, m * 48, np.float64).reshape(m, 48),
                    material=arr(d.material, n * 14, np.float64).reshape(n, 14))

    def params(self, seed=1):
        p = N.Params()
        rc = N.lib.rtcs_scene_params(self._h, int(seed), C.byref(p))
        assert rc == 0
        return p

    def camera(self, index=None, width=None, height=None):
        """Camera.InitRender(width, height) of camera `index` (default: Scene.CurrentCamera, scene size)."""
        g = self.globals()
        c = N.Camera()
        rc = N.lib.rtcs_scene_camera(self._h, g.current_camera if index is None else index,
                                     g.width if width is None else width, g.height if height is None else height, C.byref(c))
        if rc:
            raise ValueError("invalid camera index or size")
        return c

    def bvh(self):
        """Scene.Prepare (Scene.cs:39-49): (nodes pointer, n_nodes, root); pointer valid while the scene lives."""
        nodes = C.POINTER(N.BvhNode)()
        n = C.c_int32()
        root = C.c_int32()
        rc = N.lib.rtcs_scene_bvh(self._h, C.byref(nodes), C.byref(n), C.byref(root))
        assert rc == 0
        return nodes, n.value, root.value

    def bvh_array(self):
        nodes, n, root = self.bvh()
        dt = np.dtype([("bmin", "<f8", 3), ("bmax", "<f8", 3), ("left", "<i4"), ("right", "<i4"), ("prim", "<i4"), ("pad", "<i4")])
        if n == 0:
            return np.zeros(0, dtype=dt), root
        buf = (N.BvhNode * n).from_address(C.addressof(nodes.contents))
        return np.frombuffer(buf, dtype=dt).copy(), root

    def primitive_bounds(self, i):
        lo = (C.c_double * 3)()
        hi = (C.c_double * 3)()
        rc = N.lib.rtcs_scene_primitive_bounds(self._h, i, lo, hi)
        if rc:
            raise IndexError(i)
        return np.array(lo[:]), np.array(hi[:])
