"""ctypes binding of the CPU oracle (oracle/orb_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package (multi_orbslam3_b200) never does.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liborb_oracle.so")


class KeyPoint(C.Structure):
    """cv::KeyPoint binary layout (28 bytes)."""
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("size", C.c_float), ("angle", C.c_float),
                ("response", C.c_float), ("octave", C.c_int32), ("class_id", C.c_int32)]


KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                     ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")])
assert KP_DTYPE.itemsize == 28

PROJQ_DTYPE = np.dtype([("u", "<f4"), ("v", "<f4"), ("r", "<f4"), ("minl", "<i4"), ("maxl", "<i4"),
                        ("ur", "<f4"), ("angle", "<f4"), ("valid", "<i4")])
assert PROJQ_DTYPE.itemsize == 32


def build(force=False):
    src = os.path.join(_HERE, "orb_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE], stdout=sys.stderr)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        vp, i32, f32 = C.c_void_p, C.c_int, C.c_float
        L.orc_resize_linear_u8.argtypes = [vp, i32, i32, i32, vp, i32, i32, i32]
        L.orc_gaussian_blur7.argtypes = [vp, i32, i32, i32, vp, i32]
        L.orc_fast9_16.argtypes = [vp, i32, i32, i32, i32, i32, vp, i32]
        L.orc_fast9_16.restype = i32
        L.orc_fast_atan2.argtypes = [f32, f32]
        L.orc_fast_atan2.restype = f32
        L.orc_cv_round_f.argtypes = [f32]
        L.orc_cv_round_f.restype = i32
        L.orc_extractor_create.argtypes = [i32, f32, i32, i32, i32]
        L.orc_extractor_create.restype = vp
        L.orc_extractor_destroy.argtypes = [vp]
        L.orc_extractor_tables.argtypes = [vp] * 7
        L.orc_extract.argtypes = [vp, vp, i32, i32, i32, i32, i32, vp, vp, i32, vp]
        L.orc_extract.restype = i32
        L.orc_level_size.argtypes = [vp, i32, vp, vp]
        L.orc_level_image.argtypes = [vp, i32]
        L.orc_level_image.restype = vp
        L.orc_level_blurred.argtypes = [vp, i32]
        L.orc_level_blurred.restype = vp
        L.orc_level_candidates.argtypes = [vp, i32, vp]
        L.orc_level_candidates.restype = i32
        L.orc_level_keypoints.argtypes = [vp, i32, vp]
        L.orc_level_keypoints.restype = i32
        L.orc_distribute_octree.argtypes = [vp, i32, i32, i32, i32, i32, i32, vp, i32]
        L.orc_distribute_octree.restype = i32
        L.orc_hamming256.argtypes = [vp, vp]
        L.orc_hamming256.restype = i32
        L.orc_bf_knn2.argtypes = [vp, i32, vp, i32, vp, vp]
        L.orc_grid_build.argtypes = [vp, i32, f32, f32, f32, f32]
        L.orc_grid_build.restype = vp
        L.orc_grid_destroy.argtypes = [vp]
        L.orc_features_in_area.argtypes = [vp, vp, f32, f32, f32, i32, i32, vp, i32]
        L.orc_features_in_area.restype = i32
        L.orc_search_for_initialization.argtypes = [vp, vp, i32, vp, vp, i32, f32, f32, f32, f32,
                                                    vp, vp, i32, f32, i32]
        L.orc_search_for_initialization.restype = i32
        L.orc_search_by_projection.argtypes = [i32, vp, vp, i32, vp, vp, vp, i32, f32, f32, f32, f32,
                                               vp, f32, i32]
        L.orc_search_by_projection.restype = i32
        L.orc_stereo_band_match.argtypes = [vp, vp, i32, vp, vp, i32, vp, i32, f32, f32, vp, vp]
        L.orc_match_candidates.argtypes = [vp, i32, vp, vp, vp, vp, vp]
        L.orc_search_by_projection_ex.restype = i32
        L.orc_search_by_projection_ex.argtypes = [i32, vp, vp, i32, vp, vp, vp, i32, f32, f32, f32, f32, vp, f32, i32, i32, vp, C.c_double, vp, vp]
        L.orc_search_by_projection_full.restype = i32
        L.orc_search_by_projection_full.argtypes = [i32, vp, vp, i32, vp, vp, vp, i32, f32, f32, f32, f32, f32, f32, vp, f32, i32, i32, vp,
                                                    C.c_double, C.c_double, vp, vp]
        L.orc_undistort_keypoints.argtypes = [vp, i32, vp, vp, i32, vp, vp]
        L.orc_keypoints_to_msg.argtypes = [vp, i32, vp]
        L.orc_keypoints_from_msg.argtypes = [vp, i32, vp]
        L.orc_distinctive_descriptors.argtypes = [vp, vp, i32, vp]
        L.orc_search_for_triangulation.restype = i32
        L.orc_search_for_triangulation.argtypes = [vp, vp, vp, vp, i32, vp, vp, vp, i32, vp, vp, vp, vp, i32, vp, vp, vp, i32,
                                                   vp, f32, f32, vp, vp, i32, i32, i32, vp]
        L.orc_search_by_bow.restype = i32
        L.orc_search_by_bow.argtypes = [i32, vp, vp, vp, i32, vp, vp, vp, i32, vp, vp, vp, i32, vp, vp, vp, i32, f32, i32, vp]
        L.orc_vocab_create.restype = vp
        L.orc_vocab_create.argtypes = [i32, vp, vp, vp, vp, i32]
        L.orc_vocab_destroy.argtypes = [vp]
        L.orc_bow_transform_features.argtypes = [vp, vp, i32, i32, vp, vp, vp]
        L.orc_bow_transform.restype = i32
        L.orc_bow_transform.argtypes = [vp, vp, i32, i32, vp, vp, vp, vp, vp, vp]
        L.orc_compute_stereo_matches.argtypes = [vp, vp, vp, vp, i32, vp, vp, i32, f32, f32, vp, vp, vp]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _u8(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a


# ---- primitives ----
def resize_linear(src, dw, dh):
    src = _u8(src)
    dst = np.empty((dh, dw), np.uint8)
    lib().orc_resize_linear_u8(_p(src), src.shape[1], src.shape[0], src.strides[0], _p(dst), dw, dh, dw)
    return dst


def gaussian_blur7(src):
    src = _u8(src)
    dst = np.empty_like(src)
    lib().orc_gaussian_blur7(_p(src), src.shape[1], src.shape[0], src.strides[0], _p(dst), dst.strides[0])
    return dst


def fast9_16(img, threshold, nms=True):
    img = _u8(img)
    cap = img.size // 2 + 16
    out = np.empty((cap, 3), np.int32)
    n = lib().orc_fast9_16(_p(img), img.shape[1], img.shape[0], img.strides[0], threshold, int(nms), _p(out), cap)
    return out[:n].copy()


def fast_atan2(y, x):
    return lib().orc_fast_atan2(float(y), float(x))


def hamming256(a, b):
    a = _u8(a); b = _u8(b)
    return lib().orc_hamming256(_p(a), _p(b))


def bf_knn2(q, t):
    q = _u8(q); t = _u8(t)
    idx = np.empty((len(q), 2), np.int32)
    dist = np.empty((len(q), 2), np.int32)
    lib().orc_bf_knn2(_p(q), len(q), _p(t), len(t), _p(idx), _p(dist))
    return idx, dist


def distribute_octree(xyr, min_x, max_x, min_y, max_y, n_features):
    xyr = np.ascontiguousarray(xyr, np.float32).reshape(-1, 3)
    out = np.empty((len(xyr) + 8, 3), np.float32)
    n = lib().orc_distribute_octree(_p(xyr), len(xyr), min_x, max_x, min_y, max_y, n_features, _p(out), len(out))
    return out[:n].copy()


class Extractor:
    """Mirror of ORBextractor (R/orb_slam3/include/ORBextractor.h:47-113) on the oracle."""

    def __init__(self, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7):
        self.nfeatures, self.nlevels = nfeatures, nlevels
        self._h = lib().orc_extractor_create(nfeatures, scale_factor, nlevels, ini_th, min_th)
        self.scale = np.empty(nlevels, np.float32)
        self.inv_scale = np.empty(nlevels, np.float32)
        self.sigma2 = np.empty(nlevels, np.float32)
        self.inv_sigma2 = np.empty(nlevels, np.float32)
        self.features_per_level = np.empty(nlevels, np.int32)
        self.umax = np.empty(16, np.int32)
        lib().orc_extractor_tables(self._h, _p(self.scale), _p(self.inv_scale), _p(self.sigma2),
                                   _p(self.inv_sigma2), _p(self.features_per_level), _p(self.umax))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_extractor_destroy(self._h)
            self._h = None

    def __call__(self, img, lapping=(0, 0)):
        """returns (mono_index, keypoints[KP_DTYPE], descriptors[n,32])"""
        img = _u8(img)
        if img.size == 0:
            return -1, np.empty(0, KP_DTYPE), np.empty((0, 32), np.uint8)
        cap = self.nfeatures + 4 * self.nlevels + 64
        kps = np.zeros(cap, KP_DTYPE)
        desc = np.zeros((cap, 32), np.uint8)
        n = C.c_int(0)
        mono = lib().orc_extract(self._h, _p(img), img.shape[1], img.shape[0], img.strides[0],
                                 int(lapping[0]), int(lapping[1]), _p(kps), _p(desc), cap, C.byref(n))
        if mono == -2:
            raise RuntimeError("oracle output capacity too small: %d" % n.value)
        return mono, kps[:n.value].copy(), desc[:n.value].copy()

    def level_size(self, l):
        w, h = C.c_int(), C.c_int()
        lib().orc_level_size(self._h, l, C.byref(w), C.byref(h))
        return w.value, h.value

    def _img(self, ptr, l):
        if not ptr:
            return None
        w, h = self.level_size(l)
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(h, w)).copy()

    def level_image(self, l):
        return self._img(lib().orc_level_image(self._h, l), l)

    def level_blurred(self, l):
        return self._img(lib().orc_level_blurred(self._h, l), l)

    def level_candidates(self, l):
        ptr = C.c_void_p()
        n = lib().orc_level_candidates(self._h, l, C.byref(ptr))
        if n == 0:
            return np.empty((0, 3), np.float32)
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_float)), shape=(n, 3)).copy()

    def level_keypoints(self, l):
        ptr = C.c_void_p()
        n = lib().orc_level_keypoints(self._h, l, C.byref(ptr))
        if n == 0:
            return np.empty(0, KP_DTYPE)
        raw = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(n * 28,)).copy()
        return raw.view(KP_DTYPE)


def features_in_area(kps, bounds, x, y, r, min_level=-1, max_level=-1):
    kps = np.ascontiguousarray(kps, KP_DTYPE)
    g = lib().orc_grid_build(_p(kps), len(kps), *[float(b) for b in bounds])
    out = np.empty(len(kps) + 1, np.int32)
    n = lib().orc_features_in_area(g, _p(kps), float(x), float(y), float(r), min_level, max_level, _p(out), len(out))
    lib().orc_grid_destroy(g)
    return out[:n].copy()


def search_for_initialization(k1, d1, k2, d2, bounds, prev_xy, window=100, nnratio=0.9, check_ori=True):
    """bounds = (minX, maxX, minY, maxY).  Returns (nmatches, matches12, updated prev_xy)."""
    k1 = np.ascontiguousarray(k1, KP_DTYPE); k2 = np.ascontiguousarray(k2, KP_DTYPE)
    d1 = _u8(d1); d2 = _u8(d2)
    prev = np.ascontiguousarray(prev_xy, np.float32).copy()
    m12 = np.empty(len(k1), np.int32)
    n = lib().orc_search_for_initialization(_p(k1), _p(d1), len(k1), _p(k2), _p(d2), len(k2),
                                            *[float(b) for b in bounds], _p(prev), _p(m12), int(window),
                                            float(nnratio), int(check_ori))
    return n, m12, prev


def search_by_projection(mode, queries, qdesc, k2, d2, bounds, assigned=None, uright=None, nnratio=0.8, check_ori=True):
    q = np.ascontiguousarray(queries, PROJQ_DTYPE)
    qdesc = _u8(qdesc); k2 = np.ascontiguousarray(k2, KP_DTYPE); d2 = _u8(d2)
    a = np.full(len(k2), -1, np.int32) if assigned is None else np.ascontiguousarray(assigned, np.int32).copy()
    ur = None if uright is None else np.ascontiguousarray(uright, np.float32)
    n = lib().orc_search_by_projection(mode, _p(q), _p(qdesc), len(q), _p(k2), _p(d2),
                                       _p(ur) if ur is not None else None, len(k2),
                                       *[float(b) for b in bounds], _p(a), float(nnratio), int(check_ori))
    return n, a


def stereo_band_match(kl, dl, kr, dr, scale_factors, nrows, min_d, max_d):
    kl = np.ascontiguousarray(kl, KP_DTYPE); kr = np.ascontiguousarray(kr, KP_DTYPE)
    dl = _u8(dl); dr = _u8(dr)
    sf = np.ascontiguousarray(scale_factors, np.float32)
    bi = np.empty(len(kl), np.int32); bd = np.empty(len(kl), np.int32)
    lib().orc_stereo_band_match(_p(kl), _p(dl), len(kl), _p(kr), _p(dr), len(kr), _p(sf), int(nrows),
                                float(min_d), float(max_d), _p(bi), _p(bd))
    return bi, bd


def compute_stereo_matches(ex_left, ex_right, kl, dl, kr, dr, mb, mbf):
    """Frame::ComputeStereoMatches on the two oracle extractors' last pyramids -> (mvuRight, mvDepth, sad best distance)."""
    kl = np.ascontiguousarray(kl, KP_DTYPE); kr = np.ascontiguousarray(kr, KP_DTYPE)
    dl = _u8(dl); dr = _u8(dr)
    ur = np.empty(len(kl), np.float32); dp = np.empty(len(kl), np.float32); sd = np.empty(len(kl), np.int32)
    lib().orc_compute_stereo_matches(ex_left._h, ex_right._h, _p(kl), _p(dl), len(kl), _p(kr), _p(dr), len(kr),
                                     float(mb), float(mbf), _p(ur), _p(dp), _p(sd))
    return ur, dp, sd


def match_candidates(q, t, offsets, indices):
    q = _u8(q); t = _u8(t)
    off = np.ascontiguousarray(offsets, np.int32); ind = np.ascontiguousarray(indices, np.int32)
    idx = np.empty((len(q), 2), np.int32); dist = np.empty((len(q), 2), np.int32)
    lib().orc_match_candidates(_p(q), len(q), _p(t), _p(off), _p(ind), _p(idx), _p(dist))
    return idx, dist


class Vocabulary:
    """DBoW2 ORB vocabulary (node table in file order, see orb_oracle.h)."""

    def __init__(self, parent, is_leaf, desc, weight, L):
        self.parent = np.ascontiguousarray(parent, np.int32); self.is_leaf = np.ascontiguousarray(is_leaf, np.uint8)
        self.desc = _u8(desc); self.weight = np.ascontiguousarray(weight, np.float64); self.L = int(L)
        self._h = lib().orc_vocab_create(len(self.parent), _p(self.parent), _p(self.is_leaf), _p(self.desc), _p(self.weight), self.L)
        if not self._h:
            raise ValueError("invalid vocabulary")

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_vocab_destroy(self._h); self._h = None

    def transform_features(self, desc, levelsup=4):
        d = _u8(desc); n = len(d)
        w = np.empty(n, np.int32); wt = np.empty(n, np.float64); nd = np.empty(n, np.int32)
        lib().orc_bow_transform_features(self._h, _p(d), n, levelsup, _p(w), _p(wt), _p(nd))
        return w, wt, nd

    def transform(self, desc, levelsup=4):
        """-> (bow word ids, bow values), (fv node ids, list of feature-index arrays)"""
        d = _u8(desc); n = len(d)
        bw = np.empty(n + 1, np.int32); bv = np.empty(n + 1, np.float64)
        fn = np.empty(n + 1, np.int32); fs = np.empty(n + 2, np.int32); ff = np.empty(n + 1, np.int32)
        nfv = C.c_int(0)
        nb = lib().orc_bow_transform(self._h, _p(d), n, levelsup, _p(bw), _p(bv), _p(fn), _p(fs), _p(ff), C.byref(nfv))
        k = nfv.value
        return (bw[:nb].copy(), bv[:nb].copy()), (fn[:k].copy(), [ff[fs[i]:fs[i + 1]].copy() for i in range(k)])


def fv_to_csr(fv):
    """(node ids, list of feature arrays) -> (nodes int32, start int32[n+1], features int32)"""
    nodes, feats = fv
    start = np.zeros(len(nodes) + 1, np.int32)
    if len(nodes):
        start[1:] = np.cumsum([len(f) for f in feats])
    flat = np.concatenate(feats).astype(np.int32) if len(nodes) else np.zeros(0, np.int32)
    return np.ascontiguousarray(nodes, np.int32), start, np.ascontiguousarray(flat)


def search_by_bow(mode, k1, d1, valid1, fv1, k2, d2, valid2, fv2, nnratio=0.7, check_ori=True):
    k1 = np.ascontiguousarray(k1, KP_DTYPE); k2 = np.ascontiguousarray(k2, KP_DTYPE); d1 = _u8(d1); d2 = _u8(d2)
    v1 = np.ascontiguousarray(valid1, np.uint8); v2 = None if valid2 is None else np.ascontiguousarray(valid2, np.uint8)
    n1, s1, f1 = fv_to_csr(fv1); n2, s2, f2 = fv_to_csr(fv2)
    m12 = np.empty(len(k1), np.int32)
    n = lib().orc_search_by_bow(mode, _p(k1), _p(d1), _p(v1), len(k1), _p(n1), _p(s1), _p(f1), len(n1),
                                _p(k2), _p(d2), _p(v2) if v2 is not None else None, len(k2), _p(n2), _p(s2), _p(f2), len(n2),
                                float(nnratio), int(check_ori), _p(m12))
    return n, m12


def distinctive_descriptors(desc, offsets):
    d = _u8(desc); off = np.ascontiguousarray(offsets, np.int32)
    best = np.empty(len(off) - 1, np.int32)
    lib().orc_distinctive_descriptors(_p(d), _p(off), len(off) - 1, _p(best))
    return best


def undistort_keypoints(kps, K, dist, P):
    k = np.ascontiguousarray(kps, KP_DTYPE); out = np.empty_like(k)
    Kf = np.ascontiguousarray(K, np.float32).reshape(9); Pf = np.ascontiguousarray(P, np.float32).reshape(9)
    d = np.ascontiguousarray(dist, np.float32)
    lib().orc_undistort_keypoints(_p(k), len(k), _p(Kf), _p(d), len(d), _p(Pf), _p(out))
    return out


def keypoints_to_msg(kps):
    k = np.ascontiguousarray(kps, KP_DTYPE); out = np.empty((len(k), 15), np.uint8)
    lib().orc_keypoints_to_msg(_p(k), len(k), _p(out))
    return out


def keypoints_from_msg(msg):
    m = np.ascontiguousarray(msg, np.uint8).reshape(-1, 15); out = np.empty(len(m), KP_DTYPE)
    lib().orc_keypoints_from_msg(_p(m), len(m), _p(out))
    return out


def search_by_projection_full(mode, queries, qdesc, k2, d2, bounds, assigned=None, uright=None, nnratio=0.8, check_ori=True,
                              max_dist=100, inv_sigma2=None, chi2=0.0, chi2_stereo=0.0, query_origin=None):
    """orc_search_by_projection_full -> (count, assigned) for modes 0 / 1, (count, best_idx, best_dist) for mode 3"""
    q = np.ascontiguousarray(queries, PROJQ_DTYPE)
    qdesc = _u8(qdesc); k2 = np.ascontiguousarray(k2, KP_DTYPE); d2 = _u8(d2)
    a = np.full(len(k2), -1, np.int32) if assigned is None else np.ascontiguousarray(assigned, np.int32).copy()
    ur = None if uright is None else np.ascontiguousarray(uright, np.float32)
    sg = None if inv_sigma2 is None else np.ascontiguousarray(inv_sigma2, np.float32)
    bi = np.empty(len(q), np.int32); bd = np.empty(len(q), np.int32)
    qo = (bounds[0], bounds[2]) if query_origin is None else query_origin
    n = lib().orc_search_by_projection_full(mode, _p(q), _p(qdesc), len(q), _p(k2), _p(d2), _p(ur) if ur is not None else None, len(k2),
                                            *[float(b) for b in bounds], float(qo[0]), float(qo[1]), _p(a), float(nnratio), int(check_ori),
                                            int(max_dist), _p(sg) if sg is not None else None, float(chi2), float(chi2_stereo), _p(bi), _p(bd))
    return (n, bi, bd) if mode == 3 else (n, a)


def search_by_projection_rig(mode, ql, qr, qdesc, k2, d2, n_left, bounds, assigned=None, l2r=None, r2l=None, nnratio=0.8, check_ori=True,
                             max_dist=100):
    """orc_search_by_projection_rig (two-camera frame: k2 = left keypoints then right keypoints) -> (count, assigned)"""
    ql = np.ascontiguousarray(ql, PROJQ_DTYPE); qr = np.ascontiguousarray(qr, PROJQ_DTYPE)
    qdesc = _u8(qdesc); k2 = np.ascontiguousarray(k2, KP_DTYPE); d2 = _u8(d2)
    a = np.full(len(k2), -1, np.int32) if assigned is None else np.ascontiguousarray(assigned, np.int32).copy()
    pl = None if l2r is None else np.ascontiguousarray(l2r, np.int32); pr = None if r2l is None else np.ascontiguousarray(r2l, np.int32)
    L = lib()
    L.orc_search_by_projection_rig.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                               C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_float, C.c_int, C.c_int]
    n = L.orc_search_by_projection_rig(mode, _p(ql), _p(qr), _p(qdesc), len(ql), _p(k2), _p(d2), int(n_left), len(k2) - int(n_left),
                                       _p(pl) if pl is not None else None, _p(pr) if pr is not None else None,
                                       *[float(b) for b in bounds], _p(a), float(nnratio), int(check_ori), int(max_dist))
    return n, a


def search_by_projection_ex(mode, queries, qdesc, k2, d2, bounds, assigned=None, uright=None, nnratio=0.8, check_ori=True,
                            max_dist=100, inv_sigma2=None, chi2=0.0):
    """-> (count, assigned) for modes 0 / 1, (count, best_idx, best_dist) for mode 3"""
    q = np.ascontiguousarray(queries, PROJQ_DTYPE)
    qdesc = _u8(qdesc); k2 = np.ascontiguousarray(k2, KP_DTYPE); d2 = _u8(d2)
    a = np.full(len(k2), -1, np.int32) if assigned is None else np.ascontiguousarray(assigned, np.int32).copy()
    ur = None if uright is None else np.ascontiguousarray(uright, np.float32)
    sg = None if inv_sigma2 is None else np.ascontiguousarray(inv_sigma2, np.float32)
    bi = np.empty(len(q), np.int32); bd = np.empty(len(q), np.int32)
    n = lib().orc_search_by_projection_ex(mode, _p(q), _p(qdesc), len(q), _p(k2), _p(d2), _p(ur) if ur is not None else None, len(k2),
                                          *[float(b) for b in bounds], _p(a), float(nnratio), int(check_ori), int(max_dist),
                                          _p(sg) if sg is not None else None, float(chi2), _p(bi), _p(bd))
    return (n, bi, bd) if mode == 3 else (n, a)


def search_for_triangulation(k1, d1, free1, stereo1, fv1, k2, d2, free2, stereo2, fv2, F12, ep, scale2, sigma2_2,
                             only_stereo=False, coarse=False, check_ori=True):
    k1 = np.ascontiguousarray(k1, KP_DTYPE); k2 = np.ascontiguousarray(k2, KP_DTYPE); d1 = _u8(d1); d2 = _u8(d2)
    f1 = np.ascontiguousarray(free1, np.uint8); f2 = np.ascontiguousarray(free2, np.uint8)
    s1 = None if stereo1 is None else np.ascontiguousarray(stereo1, np.uint8)
    s2 = None if stereo2 is None else np.ascontiguousarray(stereo2, np.uint8)
    n1, st1, ft1 = fv_to_csr(fv1); n2, st2, ft2 = fv_to_csr(fv2)
    F = np.ascontiguousarray(F12, np.float32).reshape(9); sc = np.ascontiguousarray(scale2, np.float32); sg = np.ascontiguousarray(sigma2_2, np.float32)
    m12 = np.empty(len(k1), np.int32)
    n = lib().orc_search_for_triangulation(_p(k1), _p(d1), _p(f1), _p(s1) if s1 is not None else None, len(k1), _p(n1), _p(st1), _p(ft1), len(n1),
                                           _p(k2), _p(d2), _p(f2), _p(s2) if s2 is not None else None, len(k2), _p(n2), _p(st2), _p(ft2), len(n2),
                                           _p(F), float(ep[0]), float(ep[1]), _p(sc), _p(sg), int(only_stereo), int(coarse), int(check_ori), _p(m12))
    return n, m12
