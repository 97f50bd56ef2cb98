"""ctypes host layer over liborbx_b200.so (include/orbx.h).

Mirrors the reference's class interface for the hot path so that tests read like calls into the
reference: ORBextractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST) with operator()
(R/orb_slam3/include/ORBextractor.h:47-113) and ORBmatcher(nnratio, checkOri) with DescriptorDistance,
SearchForInitialization, SearchByProjection (R/orb_slam3/include/ORBmatcher.h:35-108).

There is no CPU fallback: importing works anywhere (so that the C ABI can be inspected), but creating an
extractor or matcher without the CUDA library or without a GPU raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liborbx_b200.so")

ORBX_OK, ORBX_E_INVALID, ORBX_E_EMPTY, ORBX_E_CUDA, ORBX_E_CAPACITY, ORBX_E_NOMEM = 0, -1, -2, -3, -4, -5
TH_HIGH, TH_LOW, HISTO_LENGTH = 100, 50, 30

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                     ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")])
PROJQ_DTYPE = np.dtype([("u", "<f4"), ("v", "<f4"), ("r", "<f4"), ("minl", "<i4"), ("maxl", "<i4"),
                        ("ur", "<f4"), ("angle", "<f4"), ("valid", "<i4")])
assert KP_DTYPE.itemsize == 28 and PROJQ_DTYPE.itemsize == 32

# every symbol include/orbx.h declares (tests check the library exports all of them)
EXPORTS = [
    "orbx_last_error", "orbx_device_count", "orbx_launch_count",
    "orbx_extractor_create", "orbx_extractor_destroy", "orbx_extractor_tables", "orbx_extractor_max_keypoints",
    "orbx_extract", "orbx_extract_batch", "orbx_extract_batch_device", "orbx_extractor_copy_slot",
    "orbx_extractor_results_device", "orbx_extractor_download", "orbx_extractor_sync", "orbx_extractor_profile",
    "orbx_pyramid_level_size", "orbx_pyramid_to_host", "orbx_pyramid_levels_to_host", "orbx_pyramid_levels_staged", "orbx_blurred_to_host", "orbx_candidates_to_host",
    "orbx_level_keypoints_to_host",
    "orbx_matcher_create", "orbx_matcher_destroy", "orbx_matcher_sync", "orbx_hamming_pairs",
    "orbx_bf_knn2", "orbx_bf_knn2_device", "orbx_knn2_merge_device",
    "orbx_search_for_initialization", "orbx_match_slots_device", "orbx_extract_match_batch", "orbx_extract_match_batch_prefetch", "orbx_stream_submit", "orbx_stream_wait", "orbx_extract_match_batch_device",
    "orbx_search_by_projection",
    "orbx_stereo_band_match", "orbx_stereo_matches", "orbx_stereo_matches_batch", "orbx_stereo_matches_batch_device",
    "orbx_extract_stereo_batch",
    "orbx_fast_segment_plan", "orbx_matcher_set_slot_keypoints", "orbx_matcher_set_camera", "orbx_matcher_undistorted_device", "orbx_search_by_projection_ex", "orbx_search_by_projection_opts", "orbx_match_candidates", "orbx_search_by_bow", "orbx_search_for_triangulation", "orbx_distinctive_descriptors", "orbx_undistort_keypoints", "orbx_undistort_slots_device",
    "orbx_keypoints_to_msg", "orbx_keypoints_from_msg", "orbx_slot_keypoints_to_msg_device", "orbx_vocab_create", "orbx_vocab_destroy", "orbx_vocab_words", "orbx_vocab_word_weights",
    "orbx_bow_transform", "orbx_bow_transform_slots_device", "orbx_popc_peak",
    "orbx_search_by_projection_rig", "orbx_search_by_bow_rig",
    "orbx_kfdb_create", "orbx_kfdb_destroy", "orbx_kfdb_ingest_msg", "orbx_kfdb_ingest_slot_device", "orbx_kfdb_size", "orbx_kfdb_device",
    "orbx_kfdb_sync", "orbx_kfdb_locate", "orbx_kfdb_knn2",
]


class OrbxError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("orbx error %d: %s" % (code, msg))
        self.code = code


class Params(C.Structure):
    _fields_ = [("nfeatures", C.c_int32), ("scale_factor", C.c_float), ("nlevels", C.c_int32),
                ("ini_th_fast", C.c_int32), ("min_th_fast", C.c_int32), ("max_width", C.c_int32),
                ("max_height", C.c_int32), ("max_batch", C.c_int32), ("device", C.c_int32),
                ("max_candidates_per_level", C.c_int32)]


class MatcherParams(C.Structure):
    _fields_ = [("device", C.c_int32), ("max_keypoints", C.c_int32), ("max_batch", C.c_int32),
                ("max_candidates", C.c_int32)]


class ProjOptions(C.Structure):
    """orbx_proj_options (include/orbx.h)"""
    _fields_ = [("bounds", C.c_float * 4), ("query_origin", C.c_float * 2), ("nnratio", C.c_float), ("check_ori", C.c_int32),
                ("max_dist", C.c_int32), ("nlevels", C.c_int32), ("inv_level_sigma2", C.c_float * 12),
                ("chi2_mono", C.c_double), ("chi2_stereo", C.c_double)]


_lib = None


def lib():
    """Loads liborbx_b200.so; raises if it has not been built (no fallback path exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise OrbxError(ORBX_E_CUDA, "liborbx_b200.so is missing: run ./build.sh (or __graft_entry__.build())")
        L = C.CDLL(LIB_PATH)
        vp, i32, f32, sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t
        L.orbx_last_error.restype = C.c_char_p
        L.orbx_device_count.restype = i32
        L.orbx_launch_count.restype = C.c_ulonglong
        L.orbx_extractor_profile.argtypes = [vp, i32, vp, vp]
        L.orbx_extract_match_batch.argtypes = [vp, vp, vp, i32, i32, i32, i32, sz, i32, i32, vp, i32, f32, i32,
                                               vp, vp, i32, vp, vp, vp, vp, vp, vp]
        L.orbx_stream_submit.argtypes = [vp, vp, vp, i32, i32, i32, i32, sz, i32, i32, vp, i32, f32, i32, vp, vp, i32, vp, vp, vp, vp, vp, vp, vp]
        L.orbx_stream_wait.argtypes = [vp, vp, C.c_longlong]
        L.orbx_extract_match_batch_prefetch.argtypes = [vp, vp, vp, i32, i32, i32, i32, sz]
        L.orbx_extract_match_batch_device.argtypes = [vp, vp, vp, i32, i32, i32, i32, sz, i32, i32, vp, i32, f32, i32,
                                                      vp, vp, vp, vp, vp]
        L.orbx_extractor_create.argtypes = [C.POINTER(Params), C.POINTER(vp)]
        L.orbx_extractor_destroy.argtypes = [vp]
        L.orbx_extractor_destroy.restype = None
        L.orbx_extractor_tables.argtypes = [vp] * 6
        L.orbx_extractor_max_keypoints.argtypes = [vp]
        L.orbx_extract.argtypes = [vp, vp, i32, i32, i32, i32, i32, vp, vp, i32, vp, vp]
        L.orbx_extract_batch.argtypes = [vp, vp, i32, i32, i32, i32, sz, i32, i32, vp, vp, i32, vp, vp]
        L.orbx_extract_batch_device.argtypes = [vp, vp, i32, i32, i32, i32, sz, i32, i32, i32, vp]
        L.orbx_extractor_copy_slot.argtypes = [vp, i32, i32, vp]
        L.orbx_extractor_results_device.argtypes = [vp] * 7
        L.orbx_extractor_download.argtypes = [vp, i32, i32, vp, vp, i32, vp, vp, vp]
        L.orbx_extractor_sync.argtypes = [vp, vp]
        L.orbx_pyramid_level_size.argtypes = [vp, i32, vp, vp]
        L.orbx_pyramid_to_host.argtypes = [vp, i32, i32, vp, i32]
        L.orbx_pyramid_levels_to_host.argtypes = [vp, i32, i32, i32, vp, vp]
        L.orbx_pyramid_levels_staged.argtypes = [vp, i32, i32, i32, vp, vp]
        L.orbx_blurred_to_host.argtypes = [vp, i32, i32, vp, i32]
        L.orbx_candidates_to_host.argtypes = [vp, i32, i32, vp, i32, vp]
        L.orbx_level_keypoints_to_host.argtypes = [vp, i32, i32, vp, i32, vp]
        L.orbx_matcher_create.argtypes = [C.POINTER(MatcherParams), C.POINTER(vp)]
        L.orbx_matcher_destroy.argtypes = [vp]
        L.orbx_matcher_destroy.restype = None
        L.orbx_matcher_sync.argtypes = [vp, vp]
        L.orbx_hamming_pairs.argtypes = [vp, vp, vp, i32, vp]
        L.orbx_bf_knn2.argtypes = [vp, vp, i32, vp, i32, vp, vp]
        L.orbx_bf_knn2_device.argtypes = [vp, vp, i32, vp, C.c_longlong, vp, vp, i32, vp]
        L.orbx_knn2_merge_device.argtypes = [vp, vp, vp, i32, i32, vp, vp, vp]
        L.orbx_search_for_initialization.argtypes = [vp, vp, vp, i32, vp, vp, i32, vp, vp, vp, i32, f32, i32, vp]
        L.orbx_match_slots_device.argtypes = [vp, vp, vp, vp, i32, vp, i32, f32, i32, vp, vp, vp, vp, vp]
        L.orbx_search_by_projection.argtypes = [vp, i32, vp, vp, i32, vp, vp, vp, i32, vp, vp, f32, i32, vp]
        L.orbx_stereo_band_match.argtypes = [vp, vp, vp, i32, vp, vp, i32, vp, i32, i32, f32, f32, vp, vp]
        L.orbx_popc_peak.argtypes = [i32, vp, vp]
        L.orbx_match_candidates.argtypes = [vp, vp, i32, vp, i32, vp, vp, vp, vp]
        L.orbx_search_by_bow.argtypes = [vp, i32, vp, vp, vp, i32, vp, vp, vp, i32, vp, vp, vp, i32, vp, vp, vp, i32, f32, i32, vp, vp]
        L.orbx_fast_segment_plan.argtypes = [i32, vp, vp, vp, vp, vp]
        L.orbx_matcher_set_slot_keypoints.argtypes = [vp, vp]
        L.orbx_matcher_set_camera.argtypes = [vp, vp, vp, i32, vp]
        L.orbx_matcher_undistorted_device.argtypes = [vp, vp]
        L.orbx_search_by_projection_ex.argtypes = [vp, i32, vp, vp, i32, vp, vp, vp, i32, vp, vp, f32, i32, i32, vp, i32, C.c_double, vp, vp, vp]
        L.orbx_search_by_projection_opts.argtypes = [vp, i32, vp, vp, i32, vp, vp, vp, i32, vp, vp, vp, vp, vp]
        L.orbx_distinctive_descriptors.argtypes = [vp, vp, vp, i32, vp]
        L.orbx_undistort_keypoints.argtypes = [vp, vp, i32, vp, vp, i32, vp, vp]
        L.orbx_keypoints_to_msg.argtypes = [vp, vp, i32, vp]
        L.orbx_keypoints_from_msg.argtypes = [vp, vp, i32, vp]
        L.orbx_slot_keypoints_to_msg_device.argtypes = [vp, i32, vp, vp]
        L.orbx_undistort_slots_device.argtypes = [vp, i32, i32, vp, vp, i32, vp, vp, vp]
        L.orbx_search_for_triangulation.argtypes = [vp, vp, vp, vp, vp, i32, vp, vp, vp, i32, vp, vp, vp, vp, i32, vp, vp, vp, i32,
                                                    vp, f32, f32, vp, vp, i32, i32, i32, i32, vp, vp]
        L.orbx_search_by_projection_rig.argtypes = [vp, i32, vp, vp, vp, i32, vp, vp, i32, i32, vp, vp, vp, vp, vp]
        L.orbx_search_by_bow_rig.argtypes = [vp, vp, vp, vp, i32, vp, vp, vp, i32, vp, vp, i32, i32, vp, vp, vp, i32, f32, i32, vp, vp, vp]
        L.orbx_kfdb_create.argtypes = [i32, C.c_longlong, i32, C.POINTER(vp)]
        L.orbx_kfdb_destroy.argtypes = [vp]; L.orbx_kfdb_destroy.restype = None
        L.orbx_kfdb_ingest_msg.argtypes = [vp, C.c_int64, vp, vp, i32, vp]
        L.orbx_kfdb_ingest_slot_device.argtypes = [vp, C.c_int64, vp, i32, vp]
        L.orbx_kfdb_size.argtypes = [vp, vp, vp, vp]
        L.orbx_kfdb_device.argtypes = [vp, vp, vp]
        L.orbx_kfdb_sync.argtypes = [vp]
        L.orbx_kfdb_locate.argtypes = [vp, vp, i32, vp, vp]
        L.orbx_kfdb_knn2.argtypes = [vp, vp, vp, i32, C.c_longlong, vp, vp]
        L.orbx_vocab_create.argtypes = [i32, i32, vp, vp, vp, vp, i32, vp]
        L.orbx_vocab_destroy.argtypes = [vp]; L.orbx_vocab_destroy.restype = None
        L.orbx_vocab_words.argtypes = [vp]
        L.orbx_vocab_word_weights.argtypes = [vp, vp, i32]
        L.orbx_bow_transform.argtypes = [vp, vp, i32, i32, vp, vp, vp]
        L.orbx_bow_transform_slots_device.argtypes = [vp, vp, i32, i32, i32, vp, vp, vp]
        L.orbx_stereo_matches.argtypes = [vp, vp, vp, i32, i32, i32, i32, f32, f32, vp, vp, vp, i32, vp]
        L.orbx_stereo_matches_batch.argtypes = [vp, vp, vp, i32, i32, f32, f32, vp, vp, i32]
        L.orbx_stereo_matches_batch_device.argtypes = [vp, vp, vp, i32, i32, f32, f32, vp, vp, vp, vp]
        L.orbx_extract_stereo_batch.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, i32, sz, f32, f32, vp, vp, vp, vp, vp, vp, i32, vp, vp]
        _lib = L
    return _lib


def _check(rc):
    if rc != ORBX_OK:
        raise OrbxError(rc, lib().orbx_last_error().decode())


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _s(stream):
    """cudaStream_t argument: None -> the handle's own stream (NULL); 0 -> the legacy default stream (torch's
    default stream has handle 0, which the C ABI spells cudaStreamLegacy = 0x1); anything else verbatim."""
    if stream is None:
        return None
    return C.c_void_p(1 if stream == 0 else stream)


def device_count():
    return lib().orbx_device_count()


class ORBextractor:
    """Drop-in mirror of ORB_SLAM3::ORBextractor (R/orb_slam3/include/ORBextractor.h:47-113)."""

    def __init__(self, nfeatures=1000, scaleFactor=1.2, nlevels=8, iniThFAST=20, minThFAST=7,
                 max_width=1280, max_height=1024, max_batch=1, device=0, max_candidates_per_level=0):
        self._h = C.c_void_p()
        self.params = Params(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, max_width, max_height,
                             max_batch, device, max_candidates_per_level)
        _check(lib().orbx_extractor_create(C.byref(self.params), C.byref(self._h)))
        self.nfeatures, self.nlevels, self.max_batch = nfeatures, nlevels, max_batch
        self.cap = lib().orbx_extractor_max_keypoints(self._h)
        t = [np.empty(nlevels, np.float32) for _ in range(4)] + [np.empty(nlevels, np.int32)]
        _check(lib().orbx_extractor_tables(self._h, *[_p(x) for x in t]))
        self.mvScaleFactor, self.mvInvScaleFactor, self.mvLevelSigma2, self.mvInvLevelSigma2, self.mnFeaturesPerLevel = t

    def close(self):
        if getattr(self, "_h", None) and self._h.value and _lib is not None:
            _lib.orbx_extractor_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    # reference getters
    def GetLevels(self): return self.nlevels
    def GetScaleFactor(self): return float(self.params.scale_factor)
    def GetScaleFactors(self): return self.mvScaleFactor
    def GetInverseScaleFactors(self): return self.mvInvScaleFactor
    def GetScaleSigmaSquares(self): return self.mvLevelSigma2
    def GetInverseScaleSigmaSquares(self): return self.mvInvLevelSigma2

    def __call__(self, image, mask=None, vLappingArea=(0, 0)):
        """operator(): returns (monoIndex, keypoints, descriptors); monoIndex is -1 for an empty image."""
        img = np.ascontiguousarray(image, np.uint8)
        if img.size == 0:
            return -1, np.empty(0, KP_DTYPE), np.empty((0, 32), np.uint8)
        assert img.ndim == 2, "CV_8UC1 expected (R/src/ORBextractor.cc:1076)"
        kps = np.zeros(self.cap, KP_DTYPE)
        desc = np.zeros((self.cap, 32), np.uint8)
        n, mono = C.c_int(0), C.c_int(0)
        rc = lib().orbx_extract(self._h, _p(img), img.shape[1], img.shape[0], img.strides[0],
                                int(vLappingArea[0]), int(vLappingArea[1]), _p(kps), _p(desc), self.cap,
                                C.byref(n), C.byref(mono))
        if rc == ORBX_E_EMPTY:
            return -1, np.empty(0, KP_DTYPE), np.empty((0, 32), np.uint8)
        _check(rc)
        return mono.value, kps[:n.value].copy(), desc[:n.value].copy()

    def extract_batch(self, images, vLappingArea=(0, 0)):
        """Host frames [B,H,W] -> list of (monoIndex, keypoints, descriptors); H2D/D2H inside."""
        imgs = np.ascontiguousarray(images, np.uint8)
        B, H, W = imgs.shape
        kps = np.zeros((B, self.cap), KP_DTYPE)
        desc = np.zeros((B, self.cap, 32), np.uint8)
        n = np.zeros(B, np.int32); mono = np.zeros(B, np.int32)
        _check(lib().orbx_extract_batch(self._h, _p(imgs), B, W, H, imgs.strides[1], imgs.strides[0],
                                        int(vLappingArea[0]), int(vLappingArea[1]), _p(kps), _p(desc), self.cap,
                                        _p(n), _p(mono)))
        return [(int(mono[i]), kps[i, :n[i]].copy(), desc[i, :n[i]].copy()) for i in range(B)]

    def extract_batch_device(self, d_ptr, batch, width, height, stride, frame_stride, vLappingArea=(0, 0),
                             first_slot=0, stream=None):
        _check(lib().orbx_extract_batch_device(self._h, C.c_void_p(d_ptr), batch, width, height, stride, frame_stride,
                                               int(vLappingArea[0]), int(vLappingArea[1]), first_slot,
                                               _s(stream)))

    def copy_slot(self, src, dst, stream=None):
        _check(lib().orbx_extractor_copy_slot(self._h, src, dst, _s(stream)))

    def sync(self, stream=None):
        _check(lib().orbx_extractor_sync(self._h, _s(stream)))

    def download(self, first_slot, count, stream=None):
        kps = np.zeros((count, self.cap), KP_DTYPE)
        desc = np.zeros((count, self.cap, 32), np.uint8)
        n = np.zeros(count, np.int32); mono = np.zeros(count, np.int32)
        _check(lib().orbx_extractor_download(self._h, first_slot, count, _p(kps), _p(desc), self.cap, _p(n), _p(mono),
                                             _s(stream)))
        return [(int(mono[i]), kps[i, :n[i]].copy(), desc[i, :n[i]].copy()) for i in range(count)]

    def profile(self, enable=-1):
        """per-stage CUDA-event milliseconds {pyramid+blur, FAST, octree, describe} summed over `batches` batches"""
        ms = (C.c_double * 4)(); nb = C.c_int()
        _check(lib().orbx_extractor_profile(self._h, enable, ms, C.byref(nb)))
        return list(ms), nb.value

    def results_device(self):
        ptrs = [C.c_void_p() for _ in range(4)]
        cap, slots = C.c_int(), C.c_int()
        _check(lib().orbx_extractor_results_device(self._h, *[C.byref(p) for p in ptrs], C.byref(cap), C.byref(slots)))
        return dict(kps=ptrs[0].value, desc=ptrs[1].value, n=ptrs[2].value, mono=ptrs[3].value, cap=cap.value, slots=slots.value)

    # mvImagePyramid and test taps
    def level_size(self, level):
        w, h = C.c_int(), C.c_int()
        _check(lib().orbx_pyramid_level_size(self._h, level, C.byref(w), C.byref(h)))
        return w.value, h.value

    def _level(self, fn, slot, level):
        w, h = self.level_size(level)
        out = np.empty((h, w), np.uint8)
        _check(fn(self._h, slot, level, _p(out), w))
        return out

    def pyramid_level(self, level, slot=0):
        return self._level(lib().orbx_pyramid_to_host, slot, level)

    def blurred_level(self, level, slot=0):
        return self._level(lib().orbx_blurred_to_host, slot, level)

    def _xyr(self, fn, slot, level, cap):
        out = np.empty((cap, 3), np.float32)
        n = C.c_int()
        _check(fn(self._h, slot, level, _p(out), cap, C.byref(n)))
        return out[:n.value].copy()

    def level_candidates(self, level, slot=0, cap=1 << 17):
        return self._xyr(lib().orbx_candidates_to_host, slot, level, cap)

    def level_keypoints(self, level, slot=0):
        return self._xyr(lib().orbx_level_keypoints_to_host, slot, level, self.cap)


class ORBmatcher:
    """Mirror of ORB_SLAM3::ORBmatcher's Hamming searches on flat arrays (R/orb_slam3/include/ORBmatcher.h:35-108)."""
    TH_LOW, TH_HIGH, HISTO_LENGTH = TH_LOW, TH_HIGH, HISTO_LENGTH

    def __init__(self, nnratio=0.6, checkOri=True, max_keypoints=8192, max_batch=1, device=0, max_candidates=0):
        self.mfNNratio, self.mbCheckOrientation = float(nnratio), bool(checkOri)
        self._h = C.c_void_p()
        self.params = MatcherParams(device, max_keypoints, max_batch, max_candidates)
        _check(lib().orbx_matcher_create(C.byref(self.params), C.byref(self._h)))
        self.K = max_keypoints

    def close(self):
        if getattr(self, "_h", None) and self._h.value and _lib is not None:
            _lib.orbx_matcher_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def sync(self, stream=None):
        _check(lib().orbx_matcher_sync(self._h, _s(stream)))

    def DescriptorDistance(self, a, b):
        """Batched ORBmatcher::DescriptorDistance: a, b are [n,32] (or [32]) uint8."""
        a = np.ascontiguousarray(a, np.uint8).reshape(-1, 32); b = np.ascontiguousarray(b, np.uint8).reshape(-1, 32)
        out = np.empty(len(a), np.int32)
        _check(lib().orbx_hamming_pairs(self._h, _p(a), _p(b), len(a), _p(out)))
        return out

    def knnMatch2(self, q, t):
        """cv::BFMatcher(NORM_HAMMING).knnMatch(q, t, k=2) -> (idx[nq,2], dist[nq,2])."""
        q = np.ascontiguousarray(q, np.uint8).reshape(-1, 32); t = np.ascontiguousarray(t, np.uint8).reshape(-1, 32)
        idx = np.full((len(q), 2), -1, np.int32); dist = np.full((len(q), 2), -1, np.int32)
        _check(lib().orbx_bf_knn2(self._h, _p(q), len(q), _p(t), len(t), _p(idx), _p(dist)))
        return idx, dist

    def SearchForInitialization(self, k1, d1, k2, d2, bounds, vbPrevMatched, windowSize=10):
        """returns (nmatches, vnMatches12, updated vbPrevMatched); bounds = (mnMinX, mnMaxX, mnMinY, mnMaxY)."""
        k1 = np.ascontiguousarray(k1, KP_DTYPE); k2 = np.ascontiguousarray(k2, KP_DTYPE)
        d1 = np.ascontiguousarray(d1, np.uint8); d2 = np.ascontiguousarray(d2, np.uint8)
        prev = np.ascontiguousarray(vbPrevMatched, np.float32).copy()
        m12 = np.full(len(k1), -1, np.int32)
        b = np.array(bounds, np.float32)
        nm = C.c_int(0)
        _check(lib().orbx_search_for_initialization(self._h, _p(k1), _p(d1), len(k1), _p(k2), _p(d2), len(k2), _p(b),
                                                    _p(prev), _p(m12), int(windowSize), self.mfNNratio,
                                                    int(self.mbCheckOrientation), C.byref(nm)))
        return nm.value, m12, prev

    def SearchByProjection(self, mode, queries, qdesc, k2, d2, bounds, assigned=None, uright=None):
        """mode 0: (Frame&, const Frame&, th, bMono); mode 1: (Frame&, vector<MapPoint*>&, th).  Returns (nmatches, assigned)."""
        q = np.ascontiguousarray(queries, PROJQ_DTYPE); qd = np.ascontiguousarray(qdesc, np.uint8)
        k2 = np.ascontiguousarray(k2, KP_DTYPE); d2 = np.ascontiguousarray(d2, np.uint8)
        a = np.full(len(k2), -1, np.int32) if assigned is None else np.ascontiguousarray(assigned, np.int32).copy()
        ur = None if uright is None else np.ascontiguousarray(uright, np.float32)
        b = np.array(bounds, np.float32)
        nm = C.c_int(0)
        _check(lib().orbx_search_by_projection(self._h, mode, _p(q), _p(qd), len(q), _p(k2), _p(d2), _p(ur), len(k2),
                                               _p(b), _p(a), self.mfNNratio, int(self.mbCheckOrientation), C.byref(nm)))
        return nm.value, a

    def StereoBandMatch(self, kl, dl, kr, dr, scale_factors, nrows, minD, maxD):
        kl = np.ascontiguousarray(kl, KP_DTYPE); kr = np.ascontiguousarray(kr, KP_DTYPE)
        dl = np.ascontiguousarray(dl, np.uint8); dr = np.ascontiguousarray(dr, np.uint8)
        sf = np.ascontiguousarray(scale_factors, np.float32)
        bi = np.full(len(kl), -1, np.int32); bd = np.full(len(kl), TH_HIGH, np.int32)
        _check(lib().orbx_stereo_band_match(self._h, _p(kl), _p(dl), len(kl), _p(kr), _p(dr), len(kr), _p(sf), len(sf),
                                            int(nrows), float(minD), float(maxD), _p(bi), _p(bd)))
        return bi, bd

    def MatchCandidates(self, q, t, offsets, indices):
        """best / second best of every query over its explicit candidate list (SearchByBoW-style inner loop)."""
        q = np.ascontiguousarray(q, np.uint8).reshape(-1, 32); t = np.ascontiguousarray(t, np.uint8).reshape(-1, 32)
        off = np.ascontiguousarray(offsets, np.int32); ind = np.ascontiguousarray(indices, np.int32)
        idx = np.full((len(q), 2), -1, np.int32); dist = np.full((len(q), 2), -1, np.int32)
        _check(lib().orbx_match_candidates(self._h, _p(q), len(q), _p(t), len(t), _p(off), _p(ind), _p(idx), _p(dist)))
        return idx, dist

    def ComputeStereoMatches(self, ex_left, ex_right, mb, mbf, slot_l=0, slot_r=0, frame_l=0, frame_r=0):
        """Frame::ComputeStereoMatches on the two extractors' last results -> (mvuRight, mvDepth, SAD distance)."""
        cap = max(ex_left.cap, ex_right.cap)
        ur = np.empty(cap, np.float32); dp = np.empty(cap, np.float32); sd = np.empty(cap, np.int32)
        n = C.c_int(0)
        _check(lib().orbx_stereo_matches(self._h, ex_left._h, ex_right._h, slot_l, slot_r, frame_l, frame_r, float(mb), float(mbf),
                                         _p(ur), _p(dp), _p(sd), cap, C.byref(n)))
        return ur[:n.value].copy(), dp[:n.value].copy(), sd[:n.value].copy()

    def SearchByProjectionOpts(self, mode, queries, qdesc, k2, d2, bounds, assigned=None, uright=None, max_dist=100, inv_sigma2=None,
                               chi2=0.0, chi2_stereo=0.0, query_origin=None):
        """orbx_search_by_projection_opts: every knob (query origin of a KeyFrame grid, both Fuse gates); same returns as below"""
        q = np.ascontiguousarray(queries, PROJQ_DTYPE); qd = np.ascontiguousarray(qdesc, np.uint8).reshape(-1, 32)
        k2 = np.ascontiguousarray(k2, KP_DTYPE); d2 = np.ascontiguousarray(d2, np.uint8).reshape(-1, 32)
        a = np.full(len(k2), -1, np.int32) if assigned is None else np.ascontiguousarray(assigned, np.int32).copy()
        ur = None if uright is None else np.ascontiguousarray(uright, np.float32)
        o = ProjOptions()
        for i in range(4):
            o.bounds[i] = float(bounds[i])
        qo = (bounds[0], bounds[2]) if query_origin is None else query_origin
        o.query_origin[0], o.query_origin[1] = float(qo[0]), float(qo[1])
        o.nnratio = self.mfNNratio; o.check_ori = int(self.mbCheckOrientation); o.max_dist = int(max_dist)
        o.nlevels = 0 if inv_sigma2 is None else len(inv_sigma2)
        for i in range(o.nlevels):
            o.inv_level_sigma2[i] = float(inv_sigma2[i])
        o.chi2_mono = float(chi2); o.chi2_stereo = float(chi2_stereo)
        nm = C.c_int(0)
        bi = np.empty(len(q), np.int32); bd = np.empty(len(q), np.int32)
        _check(lib().orbx_search_by_projection_opts(self._h, int(mode), _p(q), _p(qd), len(q), _p(k2), _p(d2), _p(ur) if ur is not None else None,
                                                    len(k2), C.byref(o), _p(a), _p(bi), _p(bd), C.byref(nm)))
        return (nm.value, bi, bd) if mode == 3 else (nm.value, a)

    def SearchByProjectionEx(self, mode, queries, qdesc, k2, d2, bounds, assigned=None, uright=None, max_dist=100, inv_sigma2=None, chi2=0.0):
        """mode 0 / 1 with an acceptance bound, or mode 3 = independent best per query (Fuse); see include/orbx.h.
        -> (count, assigned) for modes 0 / 1, (count, best_idx, best_dist) for mode 3"""
        q = np.ascontiguousarray(queries, PROJQ_DTYPE); qd = np.ascontiguousarray(qdesc, np.uint8).reshape(-1, 32)
        k2 = np.ascontiguousarray(k2, KP_DTYPE); d2 = np.ascontiguousarray(d2, np.uint8).reshape(-1, 32)
        a = np.full(len(k2), -1, np.int32) if assigned is None else np.ascontiguousarray(assigned, np.int32).copy()
        ur = None if uright is None else np.ascontiguousarray(uright, np.float32)
        sg = None if inv_sigma2 is None else np.ascontiguousarray(inv_sigma2, np.float32)
        bb = np.array(bounds, np.float32); nm = C.c_int(0)
        bi = np.empty(len(q), np.int32); bd = np.empty(len(q), np.int32)
        _check(lib().orbx_search_by_projection_ex(self._h, int(mode), _p(q), _p(qd), len(q), _p(k2), _p(d2), _p(ur) if ur is not None else None,
                                                  len(k2), _p(bb), _p(a), self.mfNNratio, int(self.mbCheckOrientation), int(max_dist),
                                                  _p(sg) if sg is not None else None, 0 if sg is None else len(sg), float(chi2),
                                                  _p(bi), _p(bd), C.byref(nm)))
        return (nm.value, bi, bd) if mode == 3 else (nm.value, a)

    def SearchByProjectionRig(self, mode, ql, qr, qdesc, k2, d2, n_left, bounds, assigned=None, l2r=None, r2l=None, max_dist=100):
        """modes 0 / 1 on a two-camera frame (orbx_search_by_projection_rig): k2 / d2 = left keypoints then right keypoints"""
        ql = np.ascontiguousarray(ql, PROJQ_DTYPE); qr = np.ascontiguousarray(qr, PROJQ_DTYPE)
        qdesc = np.ascontiguousarray(qdesc, np.uint8); k2 = np.ascontiguousarray(k2, KP_DTYPE); d2 = np.ascontiguousarray(d2, np.uint8)
        a = np.full(len(k2), -1, np.int32) if assigned is None else np.ascontiguousarray(assigned, np.int32).copy()
        pl = None if l2r is None else np.ascontiguousarray(l2r, np.int32); pr = None if r2l is None else np.ascontiguousarray(r2l, np.int32)
        o = ProjOptions()
        for i in range(4): o.bounds[i] = float(bounds[i])
        o.query_origin[0] = float(bounds[0]); o.query_origin[1] = float(bounds[2])
        o.nnratio = self.mfNNratio; o.check_ori = int(self.mbCheckOrientation); o.max_dist = int(max_dist)
        nm = C.c_int(0)
        _check(lib().orbx_search_by_projection_rig(self._h, mode, _p(ql), _p(qr), _p(qdesc), len(ql), _p(k2), _p(d2), int(n_left), len(k2) - int(n_left),
                                                   None if pl is None else _p(pl), None if pr is None else _p(pr), C.byref(o), _p(a), C.byref(nm)))
        return nm.value, a

    def SearchForTriangulation(self, k1, d1, free1, stereo1, fv1, k2, d2, free2, stereo2, fv2, F12, ep, scale2, sigma2_2,
                               only_stereo=False, coarse=False):
        """ORBmatcher::SearchForTriangulation (pinhole, one camera per keyframe) -> (nmatches, matches12)"""
        k1 = np.ascontiguousarray(k1, KP_DTYPE); k2 = np.ascontiguousarray(k2, KP_DTYPE)
        d1 = np.ascontiguousarray(d1, np.uint8).reshape(-1, 32); d2 = np.ascontiguousarray(d2, np.uint8).reshape(-1, 32)
        f1 = np.ascontiguousarray(free1, np.uint8); f2 = np.ascontiguousarray(free2, np.uint8)
        s1 = None if stereo1 is None else np.ascontiguousarray(stereo1, np.uint8)
        s2 = None if stereo2 is None else np.ascontiguousarray(stereo2, np.uint8)

        def csr(fv):
            nodes, feats = fv
            start = np.zeros(len(nodes) + 1, np.int32)
            if len(nodes):
                start[1:] = np.cumsum([len(f) for f in feats])
            flat = np.concatenate(feats).astype(np.int32) if len(nodes) else np.zeros(0, np.int32)
            return np.ascontiguousarray(nodes, np.int32), start, np.ascontiguousarray(flat)
        n1, st1, ft1 = csr(fv1); n2, st2, ft2 = csr(fv2)
        F = np.ascontiguousarray(F12, np.float32).reshape(9)
        sc = np.ascontiguousarray(scale2, np.float32); sg = np.ascontiguousarray(sigma2_2, np.float32)
        m12 = np.empty(len(k1), np.int32); nm = C.c_int(0)
        _check(lib().orbx_search_for_triangulation(self._h, _p(k1), _p(d1), _p(f1), _p(s1) if s1 is not None else None, len(k1),
                                                   _p(n1), _p(st1), _p(ft1), len(n1), _p(k2), _p(d2), _p(f2), _p(s2) if s2 is not None else None,
                                                   len(k2), _p(n2), _p(st2), _p(ft2), len(n2), _p(F), float(ep[0]), float(ep[1]), _p(sc), _p(sg),
                                                   len(sc), int(only_stereo), int(coarse), int(self.mbCheckOrientation), _p(m12), C.byref(nm)))
        return nm.value, m12

    def SearchByBoW(self, mode, k1, d1, valid1, fv1, k2, d2, valid2, fv2):
        """ORBmatcher::SearchByBoW: mode 0 = (KeyFrame, Frame), mode 1 = (KeyFrame, KeyFrame); fv = (node ids, feature lists)
        as ORBVocabulary.transform returns them.  -> (nmatches, matches12)"""
        k1 = np.ascontiguousarray(k1, KP_DTYPE); k2 = np.ascontiguousarray(k2, KP_DTYPE)
        d1 = np.ascontiguousarray(d1, np.uint8).reshape(-1, 32); d2 = np.ascontiguousarray(d2, np.uint8).reshape(-1, 32)
        v1 = np.ascontiguousarray(valid1, np.uint8); v2 = None if valid2 is None else np.ascontiguousarray(valid2, np.uint8)

        def csr(fv):
            nodes, feats = fv
            start = np.zeros(len(nodes) + 1, np.int32)
            if len(nodes):
                start[1:] = np.cumsum([len(f) for f in feats])
            flat = np.concatenate(feats).astype(np.int32) if len(nodes) else np.zeros(0, np.int32)
            return np.ascontiguousarray(nodes, np.int32), start, np.ascontiguousarray(flat)
        n1, s1, f1 = csr(fv1); n2, s2, f2 = csr(fv2)
        m12 = np.empty(len(k1), np.int32); nm = C.c_int(0)
        _check(lib().orbx_search_by_bow(self._h, int(mode), _p(k1), _p(d1), _p(v1), len(k1), _p(n1), _p(s1), _p(f1), len(n1),
                                        _p(k2), _p(d2), _p(v2) if v2 is not None else None, len(k2), _p(n2), _p(s2), _p(f2), len(n2),
                                        self.mfNNratio, int(self.mbCheckOrientation), _p(m12), C.byref(nm)))
        return nm.value, m12

    def KeyPointsToMsg(self, kps):
        """Converter::toCvKeyPointMsg for an array of keypoints -> [n, 15] bytes (ROS1 layout of CvKeyPoint.msg)"""
        k = np.ascontiguousarray(kps, KP_DTYPE); out = np.empty((len(k), 15), np.uint8)
        _check(lib().orbx_keypoints_to_msg(self._h, _p(k), len(k), _p(out)))
        return out

    def KeyPointsFromMsg(self, msg):
        mm = np.ascontiguousarray(msg, np.uint8).reshape(-1, 15); out = np.empty(len(mm), KP_DTYPE)
        _check(lib().orbx_keypoints_from_msg(self._h, _p(mm), len(mm), _p(out)))
        return out

    def UndistortKeyPoints(self, kps, K, dist, P):
        """Frame::UndistortKeyPoints: cv::undistortPoints(mvKeys, K, distCoef, I, P) -> mvKeysUn"""
        k = np.ascontiguousarray(kps, KP_DTYPE); out = np.empty_like(k)
        Kf = np.ascontiguousarray(K, np.float32).reshape(9); Pf = np.ascontiguousarray(P, np.float32).reshape(9)
        d = np.ascontiguousarray(dist, np.float32)
        _check(lib().orbx_undistort_keypoints(self._h, _p(k), len(k), _p(Kf), _p(d), len(d), _p(Pf), _p(out)))
        return out

    def DistinctiveDescriptors(self, desc, offsets):
        """MapPoint::ComputeDistinctiveDescriptors for a batch of map points (CSR runs of observed descriptors) -> best index per point"""
        d = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32); off = np.ascontiguousarray(offsets, np.int32)
        best = np.empty(len(off) - 1, np.int32)
        _check(lib().orbx_distinctive_descriptors(self._h, _p(d), _p(off), len(off) - 1, _p(best)))
        return best

    def ComputeStereoMatchesBatch(self, ex_left, ex_right, mb, mbf, first=0, count=1):
        """Frame::ComputeStereoMatches for `count` stereo pairs of the two extractors' last batch -> (mvuRight, mvDepth) [count, cap]."""
        cap = ex_left.cap
        ur = np.empty((count, cap), np.float32); dp = np.empty((count, cap), np.float32)
        _check(lib().orbx_stereo_matches_batch(self._h, ex_left._h, ex_right._h, first, count, float(mb), float(mbf), _p(ur), _p(dp), cap))
        return ur, dp

    def ExtractStereoBatch(self, ex_left, ex_right, imgs_left, imgs_right, mb, mbf, out=None):
        """host frames [B,H,W] x 2 -> dict(kps_l, desc_l, n_l, kps_r, desc_r, n_r, uright, depth), one C-ABI call"""
        il = np.ascontiguousarray(imgs_left, np.uint8); ir = np.ascontiguousarray(imgs_right, np.uint8)
        B, H, W = il.shape
        cap = ex_left.cap
        if out is None:
            out = {"kps_l": np.zeros((B, cap), KP_DTYPE), "desc_l": np.zeros((B, cap, 32), np.uint8), "n_l": np.zeros(B, np.int32),
                   "kps_r": np.zeros((B, cap), KP_DTYPE), "desc_r": np.zeros((B, cap, 32), np.uint8), "n_r": np.zeros(B, np.int32),
                   "uright": np.empty((B, cap), np.float32), "depth": np.empty((B, cap), np.float32)}
        _check(lib().orbx_extract_stereo_batch(self._h, ex_left._h, ex_right._h, _p(il), _p(ir), B, W, H, W, W * H, float(mb), float(mbf),
                                               _p(out["kps_l"]), _p(out["desc_l"]), _p(out["n_l"]), _p(out["kps_r"]), _p(out["desc_r"]),
                                               _p(out["n_r"]), cap, _p(out["uright"]), _p(out["depth"])))
        return out

    def stereo_matches_batch_device(self, ex_left, ex_right, mb, mbf, first, count, d_uright, d_depth, d_sad=None, stream=None):
        """device form: d_* are device pointers to [count][ex_left.cap] arrays; enqueued on `stream`, not synchronised."""
        _check(lib().orbx_stereo_matches_batch_device(self._h, ex_left._h, ex_right._h, first, count, float(mb), float(mbf),
                                                      C.c_void_p(d_uright), C.c_void_p(d_depth),
                                                      C.c_void_p(d_sad) if d_sad else None, _s(stream)))

    def match_slots_device(self, extractor, a, b, bounds, window, d_matches12, d_nmatches, d_knn_idx=None,
                           d_knn_dist=None, stream=None):
        """a, b: device pointers to int32 slot indices (npairs each); outputs are device pointers, row stride = max_keypoints."""
        bb = np.array(bounds, np.float32)
        _check(lib().orbx_match_slots_device(self._h, extractor._h, C.c_void_p(a[0]), C.c_void_p(b[0]), a[1], _p(bb), int(window),
                                             self.mfNNratio, int(self.mbCheckOrientation), C.c_void_p(d_matches12),
                                             C.c_void_p(d_nmatches), C.c_void_p(d_knn_idx) if d_knn_idx else None,
                                             C.c_void_p(d_knn_dist) if d_knn_dist else None,
                                             _s(stream)))


class ORBVocabulary:
    """Mirror of ORB_SLAM3::ORBVocabulary = DBoW2::TemplatedVocabulary<FORB::TDescriptor, FORB> (R/include/ORBVocabulary.h)
    for the path Frame::ComputeBoW uses: transform(features, BowVector, FeatureVector, levelsup)."""

    def __init__(self, parent, is_leaf, desc, weight, L, device=0):
        parent = np.ascontiguousarray(parent, np.int32); is_leaf = np.ascontiguousarray(is_leaf, np.uint8)
        desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32); weight = np.ascontiguousarray(weight, np.float64)
        h = C.c_void_p()
        _check(lib().orbx_vocab_create(device, len(parent), _p(parent), _p(is_leaf), _p(desc), _p(weight), int(L), C.byref(h)))
        self._h = h
        self.n_words = int(lib().orbx_vocab_words(self._h))

    def close(self):
        if getattr(self, "_h", None) and _lib is not None:
            lib().orbx_vocab_destroy(self._h)
        self._h = None

    __del__ = close

    def transform_features(self, desc, levelsup=4):
        """per-descriptor (word id, word weight, node id at level L - levelsup)"""
        d = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32); n = len(d)
        w = np.empty(n, np.int32); wt = np.empty(n, np.float64); nd = np.empty(n, np.int32)
        _check(lib().orbx_bow_transform(self._h, _p(d), n, int(levelsup), _p(w), _p(wt), _p(nd)))
        return w, wt, nd

    def transform(self, desc, levelsup=4):
        """transform(features, BowVector&, FeatureVector&, levelsup): ((word ids, L1-normalised values), (node ids, feature lists)),
        both sorted by key like the std::map the reference fills (TF_IDF weighting, L1 scoring as in ORBvoc.txt)."""
        w, wt, nd = self.transform_features(desc, levelsup)
        return assemble_bow(w, wt, nd)

    def transform_slots_device(self, extractor, first_slot, count, d_word, d_node, levelsup=4, stream=None):
        _check(lib().orbx_bow_transform_slots_device(self._h, extractor._h, first_slot, count, int(levelsup), C.c_void_p(d_word),
                                                     C.c_void_p(d_node), _s(stream)))


class _DeviceView:
    """__cuda_array_interface__ over a raw device pointer (torch.as_tensor(view, device=...) wraps it without a copy)."""

    def __init__(self, ptr, shape, typestr, owner):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}
        self._owner = owner            # keeps the handle alive as long as the view


class KeyframeDB:
    """The server's keyframe-descriptor database on one GPU (include/orbx.h, orbx_kfdb_*): KF.msg byte runs
    (R/msg/KF.msg:24-29) land in the device shard that the brute-force search and server.ShardedDescriptorDB read."""

    def __init__(self, capacity_rows, max_keyframes=65536, device=0):
        self._h = C.c_void_p()
        self.device, self.capacity = device, int(capacity_rows)
        _check(lib().orbx_kfdb_create(device, self.capacity, max_keyframes, C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            lib().orbx_kfdb_destroy(self._h); self._h = C.c_void_p()

    __del__ = close

    def ingest_msg(self, kf_id, msg_keys15, msg_desc32):
        """msg_keys15: uint8 [n, 15] (or None), msg_desc32: uint8 [n, 32], exactly the bytes of the received message."""
        d = np.ascontiguousarray(msg_desc32, np.uint8).reshape(-1, 32)
        k = None if msg_keys15 is None else np.ascontiguousarray(msg_keys15, np.uint8).reshape(-1, 15)
        assert k is None or len(k) == len(d)
        first = C.c_longlong(-1)
        _check(lib().orbx_kfdb_ingest_msg(self._h, int(kf_id), None if k is None else _p(k), _p(d), len(d), C.byref(first)))
        return first.value

    def ingest_slot(self, kf_id, extractor, slot):
        first = C.c_longlong(-1)
        _check(lib().orbx_kfdb_ingest_slot_device(self._h, int(kf_id), extractor._h, slot, C.byref(first)))
        return first.value

    def size(self):
        rows, cap, kfs = C.c_longlong(), C.c_longlong(), C.c_int()
        _check(lib().orbx_kfdb_size(self._h, C.byref(rows), C.byref(kfs), C.byref(cap)))
        return rows.value, kfs.value, cap.value

    def sync(self):
        _check(lib().orbx_kfdb_sync(self._h))

    def device_views(self):
        """(descriptors [capacity, 32] u8, keypoints [capacity, 7] f32-sized words) as __cuda_array_interface__ objects."""
        dd, dk = C.c_void_p(), C.c_void_p()
        _check(lib().orbx_kfdb_device(self._h, C.byref(dd), C.byref(dk)))
        return _DeviceView(dd.value, (self.capacity, 32), "|u1", self), _DeviceView(dk.value, (self.capacity, 7), "<i4", self)

    def locate(self, rows):
        r = np.ascontiguousarray(rows, np.int64).reshape(-1)
        kf = np.empty(len(r), np.int64); ft = np.empty(len(r), np.int32)
        _check(lib().orbx_kfdb_locate(self._h, _p(r), len(r), _p(kf), _p(ft)))
        return kf, ft

    def knn2(self, matcher, queries, idx_base=0):
        q = np.ascontiguousarray(queries, np.uint8).reshape(-1, 32)
        idx = np.empty((len(q), 2), np.int32); dist = np.empty((len(q), 2), np.int32)
        _check(lib().orbx_kfdb_knn2(self._h, matcher._h, _p(q), len(q), int(idx_base), _p(idx), _p(dist)))
        return idx, dist


def assemble_bow(word, weight, node):
    """BowVector::addWeight / normalize(L1) and FeatureVector::addFeature over per-feature (word, weight, node) arrays
    (R/Thirdparty/DBoW2/DBoW2/BowVector.cpp, FeatureVector.cpp): stopped words (weight 0) are skipped; a word's value is its
    weight added once per feature, in feature order; the L1 norm is summed in word order."""
    keep = np.nonzero(weight > 0)[0]
    order = keep[np.argsort(word[keep], kind="stable")]
    words, first, counts = np.unique(word[order], return_index=True, return_counts=True)
    vals = np.empty(len(words), np.float64)
    for i, (f0, c) in enumerate(zip(first, counts)):
        wv = weight[order[f0]]
        acc = wv
        for _ in range(c - 1):
            acc += wv
        vals[i] = acc
    norm = 0.0
    for v in vals:
        norm += abs(v)
    if norm > 0.0:
        vals = vals / norm
    order = keep[np.argsort(node[keep], kind="stable")]
    nodes, first, counts = np.unique(node[order], return_index=True, return_counts=True)
    feats = [order[f0:f0 + c].astype(np.int32) for f0, c in zip(first, counts)]
    return (words.astype(np.int32), vals), (nodes.astype(np.int32), feats)


def launch_count():
    return int(lib().orbx_launch_count())


def extract_match_batch(ex, m, imgs, lap, bounds, window, out):
    """orbx_extract_match_batch on host arrays; `out` = dict of preallocated (ideally pinned) numpy arrays
    kps[B,cap] desc[B,cap,32] n[B] mono[B] matches12[B,cap] nmatches[B] and, optionally, the BF kNN-2 tables
    knn_idx[B,cap,2] knn_dist[B,cap,2] (predecessor x frame); imgs [B,H,W] uint8."""
    B, H, W = imgs.shape
    bb = np.array(bounds, np.float32)
    _check(lib().orbx_extract_match_batch(ex._h, m._h, _p(imgs), B, W, H, imgs.strides[1], imgs.strides[0], int(lap[0]), int(lap[1]),
                                          _p(bb), int(window), m.mfNNratio, int(m.mbCheckOrientation), _p(out["kps"]), _p(out["desc"]),
                                          ex.cap, _p(out["n"]), _p(out["mono"]), _p(out["matches12"]), _p(out["nmatches"]),
                                          _p(out["knn_idx"]) if "knn_idx" in out else None,
                                          _p(out["knn_dist"]) if "knn_dist" in out else None))


def stream_submit(ex, m, imgs, lap, bounds, window, out):
    """orbx_stream_submit: queue one batch (pinned [B,H,W] frames, pinned result arrays as for extract_match_batch) -> ticket"""
    B, H, W = imgs.shape
    bb = np.array(bounds, np.float32)
    t = C.c_longlong(-1)
    _check(lib().orbx_stream_submit(ex._h, m._h, _p(imgs), B, W, H, imgs.strides[1], imgs.strides[0], int(lap[0]), int(lap[1]),
                                    _p(bb), int(window), m.mfNNratio, int(m.mbCheckOrientation), _p(out["kps"]), _p(out["desc"]),
                                    ex.cap, _p(out["n"]), _p(out["mono"]), _p(out["matches12"]), _p(out["nmatches"]),
                                    _p(out["knn_idx"]) if "knn_idx" in out else None,
                                    _p(out["knn_dist"]) if "knn_dist" in out else None, C.byref(t)))
    return t.value


def stream_wait(ex, m, ticket):
    _check(lib().orbx_stream_wait(ex._h, m._h, int(ticket)))


def extract_match_batch_prefetch(ex, m, imgs):
    """orbx_extract_match_batch_prefetch: start the H2D copy of the next extract_match_batch call's frames (pinned [B,H,W] uint8)."""
    B, H, W = imgs.shape
    _check(lib().orbx_extract_match_batch_prefetch(ex._h, m._h, _p(imgs), B, W, H, imgs.strides[1], imgs.strides[0]))


def extract_match_batch_device(ex, m, d_ptr, batch, width, height, stride, frame_stride, lap, bounds, window,
                               d_matches12, d_nmatches, d_knn_idx=None, d_knn_dist=None, stream=None):
    """orbx_extract_match_batch_device: one asynchronous extract + match step on device-resident frames."""
    bb = np.array(bounds, np.float32)
    _check(lib().orbx_extract_match_batch_device(ex._h, m._h, C.c_void_p(d_ptr), batch, width, height, stride, frame_stride,
                                                 int(lap[0]), int(lap[1]), _p(bb), int(window), m.mfNNratio, int(m.mbCheckOrientation),
                                                 C.c_void_p(d_matches12), C.c_void_p(d_nmatches),
                                                 C.c_void_p(d_knn_idx) if d_knn_idx else None,
                                                 C.c_void_p(d_knn_dist) if d_knn_dist else None, _s(stream)))


def popc_peak(device=0):
    p, l = C.c_double(), C.c_double()
    _check(lib().orbx_popc_peak(device, C.byref(p), C.byref(l)))
    return p.value, l.value


def bind_to_gpu_numa(device=0):
    """Pins the calling process to the CPUs of the NUMA node GPU `device` hangs off, so that pinned host buffers allocated
    afterwards (first touch) and the threads that feed them are local to that GPU's PCIe root.  With one process per GPU this keeps
    eight concurrent host<->device streams off the inter-socket link.  Returns {"pci", "node", "cpus"}; a box without NUMA
    information (one node, sysfs absent) is left untouched."""
    info = {"pci": None, "node": None, "cpus": None}
    try:
        rt = C.CDLL("libcudart.so.12")
        buf = C.create_string_buffer(32)
        if rt.cudaDeviceGetPCIBusId(buf, C.c_int(32), C.c_int(device)) != 0:
            return info
        pci = buf.value.decode().lower()
        info["pci"] = pci
        with open("/sys/bus/pci/devices/%s/numa_node" % pci) as f:
            node = int(f.read().strip())
        info["node"] = node
        if node < 0:
            return info
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            cpulist = f.read().strip()
        cpus = set()
        for part in cpulist.split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if allowed:
            os.sched_setaffinity(0, allowed)
            info["cpus"] = "%d cpus of node %d" % (len(allowed), node)
    except (OSError, ValueError, AttributeError):
        pass
    return info
