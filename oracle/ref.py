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
import sys

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
        subprocess.check_call(["make", "-s", "-C", os.path.join(_HERE, "ref")], stdout=sys.stderr)
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
    vp, i32, f32 = C.c_void_p, C.c_int, C.c_float
    L.ref_hamming256.argtypes = [vp, vp]
    L.ref_hamming256.restype = i32
    L.ref_features_in_area.argtypes = [vp, i32, f32, f32, f32, f32, f32, f32, f32, i32, i32, vp, i32]
    L.ref_features_in_area.restype = i32
    L.ref_search_for_initialization.argtypes = [vp, vp, i32, vp, vp, i32, f32, f32, f32, f32, vp, vp, i32, f32, i32]
    L.ref_search_for_initialization.restype = i32
    L.ref_compute_stereo_matches.argtypes = [vp, vp, vp, vp, i32, vp, vp, i32, vp, i32, f32, f32, vp, vp]
    L.ref_vocab_create.argtypes = [i32, vp, vp, vp, vp, i32, i32]
    L.ref_vocab_create.restype = vp
    L.ref_vocab_destroy.argtypes = [vp]
    L.ref_bow_transform.argtypes = [vp, vp, i32, i32, vp, vp, vp, vp, vp, vp]
    L.ref_bow_transform.restype = i32
    L.ref_bow_transform_features.argtypes = [vp, vp, i32, i32, vp, vp, vp]
    L.ref_search_by_bow.argtypes = [i32, vp, vp, vp, i32, vp, vp, vp, i32, vp, vp, vp, i32, vp, vp, vp, i32, f32, i32, vp]
    L.ref_search_by_bow.restype = i32
    L.ref_distinctive_descriptors.argtypes = [vp, vp, i32, vp]


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


# ---- matcher side: same call shapes as oracle.oracle ----
def hamming256(a, b):
    a = _u8(a); b = _u8(b)
    return lib().ref_hamming256(_p(a), _p(b))


def features_in_area(kps, bounds, x, y, r, min_level=-1, max_level=-1):
    kps = np.ascontiguousarray(kps, KP_DTYPE)
    out = np.empty(len(kps) + 1, np.int32)
    n = lib().ref_features_in_area(_p(kps), len(kps), *[float(b) for b in bounds], float(x), float(y), float(r), min_level, max_level,
                                   _p(out), len(out))
    return out[:n].copy()


def search_for_initialization(k1, d1, k2, d2, bounds, prev_xy, window=100, nnratio=0.9, check_ori=True):
    k1 = np.ascontiguousarray(k1, KP_DTYPE); k2 = np.ascontiguousarray(k2, KP_DTYPE)
    d1 = _u8(d1); d2 = _u8(d2)
    prev = np.ascontiguousarray(prev_xy, np.float32).copy()
    m12 = np.empty(len(k1), np.int32)
    n = lib().ref_search_for_initialization(_p(k1), _p(d1), len(k1), _p(k2), _p(d2), len(k2), *[float(b) for b in bounds], _p(prev), _p(m12),
                                            int(window), float(nnratio), int(check_ori))
    return n, m12, prev


def compute_stereo_matches(ex_left, ex_right, kl, dl, kr, dr, mb, mbf):
    """Frame::ComputeStereoMatches on two reference extractors' last pyramids -> (mvuRight, mvDepth)"""
    kl = np.ascontiguousarray(kl, KP_DTYPE); kr = np.ascontiguousarray(kr, KP_DTYPE)
    dl = _u8(dl); dr = _u8(dr)
    ur = np.empty(len(kl), np.float32); dp = np.empty(len(kl), np.float32)
    sc = np.ascontiguousarray(ex_left.scale, np.float32)
    lib().ref_compute_stereo_matches(ex_left._h, ex_right._h, _p(kl), _p(dl), len(kl), _p(kr), _p(dr), len(kr), _p(sc), len(sc),
                                     float(mb), float(mbf), _p(ur), _p(dp))
    return ur, dp


class Vocabulary:
    """DBoW2::TemplatedVocabulary<FORB> of the reference, loaded from a node table through its own text loader."""

    def __init__(self, parent, is_leaf, desc, weight, L, k=10):
        parent = np.ascontiguousarray(parent, np.int32); is_leaf = np.ascontiguousarray(is_leaf, np.uint8)
        desc = _u8(desc); weight = np.ascontiguousarray(weight, np.float64)
        self._h = lib().ref_vocab_create(len(parent), _p(parent), _p(is_leaf), _p(desc), _p(weight), int(k), int(L))
        if not self._h:
            raise ValueError("the reference's loader rejected the vocabulary")

    def __del__(self):
        if getattr(self, "_h", None):
            lib().ref_vocab_destroy(self._h); self._h = None

    def transform_features(self, desc, levelsup=4):
        d = _u8(desc); n = len(d)
        w = np.empty(n, np.int32); wt = np.empty(n, np.float64); nd = np.empty(n, np.int32)
        lib().ref_bow_transform_features(self._h, _p(d), n, levelsup, _p(w), _p(wt), _p(nd))
        return w, wt, nd

    def transform(self, desc, levelsup=4):
        d = _u8(desc); n = len(d)
        bw = np.empty(n + 1, np.int32); bv = np.empty(n + 1, np.float64)
        fn = np.empty(n + 1, np.int32); fs = np.empty(n + 2, np.int32); ff = np.empty(n + 1, np.int32)
        nfv = C.c_int(0)
        nb = lib().ref_bow_transform(self._h, _p(d), n, levelsup, _p(bw), _p(bv), _p(fn), _p(fs), _p(ff), C.byref(nfv))
        k = nfv.value
        return (bw[:nb].copy(), bv[:nb].copy()), (fn[:k].copy(), [ff[fs[i]:fs[i + 1]].copy() for i in range(k)])


def search_by_bow(mode, k1, d1, valid1, fv1, k2, d2, valid2, fv2, nnratio=0.7, check_ori=True):
    from .oracle import fv_to_csr
    k1 = np.ascontiguousarray(k1, KP_DTYPE); k2 = np.ascontiguousarray(k2, KP_DTYPE); d1 = _u8(d1); d2 = _u8(d2)
    v1 = np.ascontiguousarray(valid1, np.uint8); v2 = None if valid2 is None else np.ascontiguousarray(valid2, np.uint8)
    n1, s1, f1 = fv_to_csr(fv1); n2, s2, f2 = fv_to_csr(fv2)
    m12 = np.empty(len(k1), np.int32)
    n = lib().ref_search_by_bow(mode, _p(k1), _p(d1), _p(v1), len(k1), _p(n1), _p(s1), _p(f1), len(n1),
                                _p(k2), _p(d2), _p(v2) if v2 is not None else None, len(k2), _p(n2), _p(s2), _p(f2), len(n2),
                                float(nnratio), int(check_ori), _p(m12))
    return n, m12


def distinctive_descriptors(desc, offsets):
    d = _u8(desc); off = np.ascontiguousarray(offsets, np.int32)
    best = np.empty(len(off) - 1, np.int32)
    lib().ref_distinctive_descriptors(_p(d), _p(off), len(off) - 1, _p(best))
    return best
