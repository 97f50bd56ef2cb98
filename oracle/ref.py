"""ctypes binding of oracle/_ref/libref_orb.so: the reference's OWN sources (R/src/ORBextractor.cc, R/src/ORBmatcher.cc,
DBoW2, Pinhole, function-level extracts of Frame / KeyFrame / MapPoint), compiled unmodified by oracle/ref/Makefile against the
minimal OpenCV stand-in dropin/cvmin.  TEST INFRASTRUCTURE ONLY: tests/, __graft_entry__.smoke() and bench.py's CPU legs may
import this module; the product package never does.

/root/reference exists only in the build container: there `build()` (re)compiles oracle/_ref; on the GPU box the prebuilt
libraries that travelled with the snapshot are used as they are.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from .oracle import KP_DTYPE, _p, _u8

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(_HERE, "_ref")
REFROOT = "/root/reference/src/orb_slam3_ros/orb_slam3"


def reference_present():
    return os.path.exists(os.path.join(REFROOT, "src", "ORBextractor.cc"))


def available(variant="arena"):
    return os.path.exists(_path(variant)) or reference_present()


def _path(variant):
    return os.path.join(REF_DIR, "libref_orb.so" if variant == "arena" else "libref_orb_malloc.so")


def build():
    """make -C oracle/ref when the reference sources are here; otherwise the prebuilt oracle/_ref is used."""
    if reference_present():
        subprocess.check_call(["make", "-s", "-C", os.path.join(_HERE, "ref")])
    return os.path.exists(_path("arena"))


_libs = {}


def lib(variant="arena"):
    if variant not in _libs:
        build()
        if not os.path.exists(_path(variant)):
            raise RuntimeError("oracle/_ref is not built and /root/reference is absent")
        L = C.CDLL(_path(variant))
        vp, i32, f32 = C.c_void_p, C.c_int, C.c_float
        L.ref_uses_arena.restype = i32
        L.ref_extractor_create.argtypes = [i32, f32, i32, i32, i32]
        L.ref_extractor_create.restype = vp
        L.ref_extractor_destroy.argtypes = [vp]
        L.ref_extractor_tables.argtypes = [vp] * 7
        L.ref_extract.argtypes = [vp, vp, i32, i32, i32, i32, i32, vp, vp, i32, vp]
        L.ref_extract.restype = i32
        L.ref_level_size.argtypes = [vp, i32, vp, vp]
        L.ref_level_image.argtypes = [vp, i32, vp, i32]
        L.ref_octree_keypoints.argtypes = [vp, vp, i32, i32, i32, i32, vp, i32]
        L.ref_octree_keypoints.restype = i32
        L.ref_distribute_octree.argtypes = [vp, vp, i32, i32, i32, i32, i32, i32, i32, vp, i32]
        L.ref_distribute_octree.restype = i32
        _bind_matcher(L)
        _libs[variant] = L
    return _libs[variant]


def _bind_matcher(L):
    pass


class Extractor:
    """ORB_SLAM3::ORBextractor of the reference itself (R/include/ORBextractor.h:47-113)."""

    def __init__(self, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7, variant="arena"):
        self.nfeatures, self.nlevels = nfeatures, nlevels
        self._L = lib(variant)
        self._h = self._L.ref_extractor_create(nfeatures, scale_factor, nlevels, ini_th, min_th)
        self.scale = np.empty(nlevels, np.float32)
        self.inv_scale = np.empty(nlevels, np.float32)
        self.sigma2 = np.empty(nlevels, np.float32)
        self.inv_sigma2 = np.empty(nlevels, np.float32)
        self.features_per_level = np.empty(nlevels, np.int32)
        self.umax = np.empty(16, np.int32)
        self._L.ref_extractor_tables(self._h, _p(self.scale), _p(self.inv_scale), _p(self.sigma2), _p(self.inv_sigma2),
                                     _p(self.features_per_level), _p(self.umax))

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.ref_extractor_destroy(self._h)
            self._h = None

    def __call__(self, img, lapping=(0, 0)):
        """operator(): returns (mono_index, keypoints[KP_DTYPE], descriptors[n, 32])"""
        img = _u8(img)
        if img.size == 0:
            return self._L.ref_extract(self._h, None, 0, 0, 0, 0, 0, None, None, 0, None), np.empty(0, KP_DTYPE), np.empty((0, 32), np.uint8)
        cap = self.nfeatures + 4 * self.nlevels + 64
        kps = np.zeros(cap, KP_DTYPE)
        desc = np.zeros((cap, 32), np.uint8)
        n = C.c_int(0)
        mono = self._L.ref_extract(self._h, _p(img), img.shape[1], img.shape[0], img.strides[0], int(lapping[0]), int(lapping[1]),
                                   _p(kps), _p(desc), cap, C.byref(n))
        assert n.value <= cap
        return mono, kps[:n.value].copy(), desc[:n.value].copy()

    def level_size(self, l):
        w, h = C.c_int(), C.c_int()
        if self._L.ref_level_size(self._h, l, C.byref(w), C.byref(h)) != 0:
            return None
        return w.value, h.value

    def level_image(self, l):
        w, h = self.level_size(l)
        out = np.empty((h, w), np.uint8)
        self._L.ref_level_image(self._h, l, _p(out), w)
        return out

    def octree_keypoints(self, img, level):
        """allKeypoints[level] of ComputeKeyPointsOctTree (before scaling)"""
        img = _u8(img)
        cap = self.nfeatures + 64
        kps = np.zeros(cap, KP_DTYPE)
        n = self._L.ref_octree_keypoints(self._h, _p(img), img.shape[1], img.shape[0], img.strides[0], level, _p(kps), cap)
        return kps[:n].copy()

    def distribute_octree(self, xyr, min_x, max_x, min_y, max_y, n_features, level=0):
        xyr = np.ascontiguousarray(xyr, np.float32).reshape(-1, 3)
        out = np.empty((len(xyr) + 8, 3), np.float32)
        n = self._L.ref_distribute_octree(self._h, _p(xyr), len(xyr), min_x, max_x, min_y, max_y, n_features, level, _p(out), len(out))
        return out[:n].copy()
