"""Generates tests/golden/*.npz: golden vectors for the ORB front-end hot path.  TEST INFRASTRUCTURE.

The reference ships no golden vectors or tests for this path and cannot be built
here, so the fixtures are produced by a SECOND, independent restatement of the reference written in
Python on top of the reference's real third-party dependency (OpenCV, here cv2 4.13.0: cv2.resize,
cv2.FastFeatureDetector, cv2.GaussianBlur, cv2.fastAtan2, cv2.BFMatcher) and cross-checked against the
C oracle (oracle/orb_oracle.c) before being written.  Run from the repo root:

    python oracle/gen_golden.py

R/ = /root/reference/src/orb_slam3_ros/orb_slam3/ ; line numbers cite R/src/ORBextractor.cc.
Needs cv2; the committed fixtures do not.
"""
import hashlib
import math
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cv2  # noqa: E402

from multi_orbslam3_b200 import synth  # noqa: E402
from oracle import oracle as O  # noqa: E402

cv2.setNumThreads(1)
f32 = np.float32
PATTERN = np.array([int(v) for v in "".join(
    l for l in open(os.path.join(os.path.dirname(__file__), "..", "multi_orbslam3_b200", "csrc", "orb_pattern.inc"))
    if not l.lstrip().startswith(("/*", "*"))).replace("\n", "").split(",") if v.strip()], np.int32).reshape(512, 2)


def cv_round(v):
    return int(np.rint(v))  # half-to-even


class PyExtractor:
    """Literal Python restatement of ORBextractor (:408-468, :763-878, :1068-1177) over cv2 primitives."""

    def __init__(self, nfeatures, scale_factor, nlevels, ini_th, min_th):
        self.nfeatures, self.nlevels, self.ini_th, self.min_th = nfeatures, nlevels, ini_th, min_th
        sf = float(f32(scale_factor))                       # double member initialised from a float (:410)
        self.scale = [f32(1.0)]
        for i in range(1, nlevels):
            self.scale.append(f32(float(self.scale[-1]) * sf))
        self.inv_scale = [f32(1.0) / s for s in self.scale]
        factor = f32(1.0 / sf)
        nd = f32(f32(nfeatures) * (f32(1) - factor)) / (f32(1) - f32(float(factor) ** nlevels))
        self.fpl, tot = [], 0
        for _ in range(nlevels - 1):
            self.fpl.append(cv_round(nd)); tot += self.fpl[-1]; nd = f32(nd * factor)
        self.fpl.append(max(nfeatures - tot, 0))
        hp = 15
        umax = [0] * (hp + 2)
        vmax = int(math.floor(hp * math.sqrt(2.0) / 2 + 1)); vmin = int(math.ceil(hp * math.sqrt(2.0) / 2))
        for v in range(vmax + 1):
            umax[v] = cv_round(math.sqrt(hp * hp - v * v))
        v0 = 0
        for v in range(hp, vmin - 1, -1):
            while umax[v0] == umax[v0 + 1]:
                v0 += 1
            umax[v] = v0; v0 += 1
        self.umax = umax[:hp + 1]
        self.fast = cv2.FastFeatureDetector_create(threshold=ini_th, nonmaxSuppression=True,
                                                   type=cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)

    def pyramid(self, img):
        lv = []
        for l in range(self.nlevels):
            w = cv_round(f32(img.shape[1]) * self.inv_scale[l]); h = cv_round(f32(img.shape[0]) * self.inv_scale[l])
            lv.append(img.copy() if l == 0 else cv2.resize(lv[l - 1], (w, h), interpolation=cv2.INTER_LINEAR))
        return lv

    def _fast(self, sub, th):
        self.fast.setThreshold(th)
        return [(kp.pt[0], kp.pt[1], kp.response) for kp in self.fast.detect(np.ascontiguousarray(sub))]

    def candidates(self, im):
        minb = 16; maxbx = im.shape[1] - 16; maxby = im.shape[0] - 16
        width = f32(maxbx - minb); height = f32(maxby - minb)
        ncols = int(width / f32(30)); nrows = int(height / f32(30))
        if ncols <= 0 or nrows <= 0:      # the reference would divide by zero (:784-785); guard as the C oracle does
            return [], (minb, maxbx, minb, maxby)
        wcell = int(math.ceil(width / f32(ncols))); hcell = int(math.ceil(height / f32(nrows)))
        out = []
        for i in range(nrows):
            iny = minb + i * hcell; mxy = iny + hcell + 6
            if iny >= maxby - 3:
                continue
            mxy = min(mxy, maxby)
            for j in range(ncols):
                inx = minb + j * wcell; mxx = inx + wcell + 6
                if inx >= maxbx - 6:
                    continue
                mxx = min(mxx, maxbx)
                sub = im[iny:mxy, inx:mxx]
                ks = self._fast(sub, self.ini_th)
                if not ks:
                    ks = self._fast(sub, self.min_th)
                out += [(x + j * wcell, y + i * hcell, r) for (x, y, r) in ks]
        return out, (minb, maxbx, minb, maxby)

    @staticmethod
    def octree(pts, minx, maxx, miny, maxy, N):
        """DistributeOctTree (:537-761); node = dict; list order kept in a Python list (front = index 0).
        Canonical tie rule for the sort at :682: equal sizes -> later-created node first."""
        seq = [0]

        def node(x0, y0, x1, y1):
            seq[0] += 1
            return {"b": (x0, y0, x1, y1), "k": [], "nomore": False, "seq": seq[0]}

        def divide(p):
            x0, y0, x1, y1 = p["b"]
            hx = int(math.ceil(f32(x1 - x0) / 2)); hy = int(math.ceil(f32(y1 - y0) / 2))
            c = [node(x0, y0, x0 + hx, y0 + hy), node(x0 + hx, y0, x1, y0 + hy),
                 node(x0, y0 + hy, x0 + hx, y1), node(x0 + hx, y0 + hy, x1, y1)]
            for i in p["k"]:
                x, y, _ = pts[i]
                if x < x0 + hx:
                    (c[0] if y < y0 + hy else c[2])["k"].append(i)
                else:
                    (c[1] if y < y0 + hy else c[3])["k"].append(i)
            for n in c:
                if len(n["k"]) == 1:
                    n["nomore"] = True
            return c

        nini = int(math.floor(float(f32(maxx - minx) / f32(maxy - miny)) + 0.5))
        hX = f32(maxx - minx) / f32(nini)
        roots = [node(int(hX * f32(i)), 0, int(hX * f32(i + 1)), maxy - miny) for i in range(nini)]
        for i, (x, _, _) in enumerate(pts):
            roots[int(f32(x) / hX)]["k"].append(i)
        L = []
        for r in roots:
            if len(r["k"]) == 1:
                r["nomore"] = True
            if r["k"]:
                L.append(r)
        finish = False
        while not finish:
            prev = len(L)
            newfront, keep, vec = [], [], []
            for n in L:
                if n["nomore"]:
                    keep.append(n); continue
                for c in divide(n):
                    if c["k"]:
                        newfront.insert(0, c)
                        if len(c["k"]) > 1:
                            vec.append(c)
            L = newfront + keep
            ntoexp = len(vec)
            if len(L) >= N or len(L) == prev:
                finish = True
            elif len(L) + 3 * ntoexp > N:
                while not finish:
                    prev = len(L)
                    pv = sorted(vec, key=lambda n: (len(n["k"]), n["seq"]))
                    vec = []
                    for n in reversed(pv):
                        for c in divide(n):
                            if c["k"]:
                                L.insert(0, c)
                                if len(c["k"]) > 1:
                                    vec.append(c)
                        L.remove(n)
                        if len(L) >= N:
                            break
                    if len(L) >= N or len(L) == prev:
                        finish = True
        res = []
        for n in L:
            best = n["k"][0]
            for i in n["k"][1:]:
                if pts[i][2] > pts[best][2]:
                    best = i
            res.append(pts[best])
        return res

    def ic_angle(self, im, x, y):
        m01 = m10 = 0
        cx, cy = cv_round(x), cv_round(y)
        for u in range(-15, 16):
            m10 += u * int(im[cy, cx + u])
        for v in range(1, 16):
            d = self.umax[v]; vs = 0
            for u in range(-d, d + 1):
                p, m = int(im[cy + v, cx + u]), int(im[cy - v, cx + u])
                vs += p - m; m10 += u * (p + m)
            m01 += v * vs
        return cv2.fastAtan2(float(m01), float(m10))

    @staticmethod
    def descriptor(blur, x, y, angle):
        ang = f32(angle) * f32(math.pi / 180.0)
        # `cos(angle)` on a float under `using namespace std` is the float overload (:65, :111);
        # glibc cosf/sinf are correctly rounded for these inputs, emulate by rounding the double result
        a = f32(math.cos(float(ang))); b = f32(math.sin(float(ang)))
        cx, cy = cv_round(x), cv_round(y)
        px = PATTERN[:, 0].astype(f32); py = PATTERN[:, 1].astype(f32)
        ry = np.rint((px * b).astype(f32) + (py * a).astype(f32)).astype(np.int64)
        rx = np.rint((px * a).astype(f32) - (py * b).astype(f32)).astype(np.int64)
        v = blur[cy + ry, cx + rx].astype(np.int32)
        bits = (v[0::2] < v[1::2]).astype(np.uint8)
        return np.packbits(bits, bitorder="little")

    def __call__(self, img, lapping):
        lv = self.pyramid(img)
        per_level, cands = [], []
        for l, im in enumerate(lv):
            c, (a, b, c0, d) = self.candidates(im)
            cands.append(np.array(c, f32).reshape(-1, 3))
            kept = self.octree(c, a, b, c0, d, self.fpl[l]) if c else []
            size = float(int(f32(31) * self.scale[l]))
            per_level.append([[x + 16, y + 16, size, self.ic_angle(im, x + 16, y + 16), r, l] for (x, y, r) in kept])
        n = sum(len(p) for p in per_level)
        kps = np.zeros(n, O.KP_DTYPE); desc = np.zeros((n, 32), np.uint8)
        mono, stereo = 0, n - 1
        blurred = []
        for l, kl in enumerate(per_level):
            if not kl:
                blurred.append(None); continue
            bl = cv2.GaussianBlur(lv[l], (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
            blurred.append(bl)
            for (x, y, size, ang, r, octv) in kl:
                d = self.descriptor(bl, x, y, ang)
                fx, fy = f32(x), f32(y)
                if l != 0:
                    fx, fy = f32(fx * self.scale[l]), f32(fy * self.scale[l])
                if lapping[0] <= fx <= lapping[1]:
                    slot = stereo; stereo -= 1
                else:
                    slot = mono; mono += 1
                kps[slot] = (fx, fy, size, ang, r, octv, -1)
                desc[slot] = d
        return mono, kps, desc, lv, blurred, cands


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def golden_extract(name, img, params, lapping, outdir):
    pe = PyExtractor(*params)
    mono, kps, desc, lv, blurred, cands = pe(img, lapping)
    ce = O.Extractor(*params)
    cmono, ckps, cdesc = ce(img, lapping)
    assert mono == cmono and len(kps) == len(ckps), (name, mono, cmono, len(kps), len(ckps))
    assert kps.tobytes() == ckps.tobytes(), name + ": keypoints differ between the two restatements"
    ndiff = int(np.count_nonzero((desc != cdesc).any(axis=1)))
    assert ndiff == 0, "%s: %d descriptors differ" % (name, ndiff)
    for l in range(params[2]):
        assert np.array_equal(lv[l], ce.level_image(l)), (name, "level", l)
        assert np.array_equal(cands[l], ce.level_candidates(l)), (name, "cands", l)
        if blurred[l] is not None:
            assert np.array_equal(blurred[l], ce.level_blurred(l)), (name, "blur", l)
    np.savez_compressed(
        os.path.join(outdir, name + ".npz"), image=img, params=np.array(params, np.float64),
        lapping=np.array(lapping, np.int32), mono_index=np.int32(mono), keypoints=kps, descriptors=desc,
        level_sha=np.array([sha(x) for x in lv]),
        blur_sha=np.array([sha(x) if x is not None else "" for x in blurred]),
        ncand=np.array([len(c) for c in cands], np.int32),
        cand_sha=np.array([sha(c) for c in cands]),
        level_sizes=np.array([[x.shape[1], x.shape[0]] for x in lv], np.int32))
    print("%-28s n=%d mono=%d cands=%s" % (name, len(kps), mono, [len(c) for c in cands]))
    return kps, desc


def golden_primitives(outdir):
    rng = np.random.default_rng(7)
    img = synth.rects_frame(200, 150, 3)
    rs = {}
    for (dw, dh) in ((167, 125), (139, 104), (100, 75), (199, 149), (260, 190)):
        rs["resize_%dx%d" % (dw, dh)] = cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR)
        assert np.array_equal(rs["resize_%dx%d" % (dw, dh)], O.resize_linear(img, dw, dh))
    blur = cv2.GaussianBlur(img, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
    assert np.array_equal(blur, O.gaussian_blur7(img))
    fd = cv2.FastFeatureDetector_create(threshold=20, nonmaxSuppression=True, type=cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    fast = {}
    for th in (7, 20, 40):
        fd.setThreshold(th)
        k = np.array([[int(p.pt[0]), int(p.pt[1]), int(p.response)] for p in fd.detect(img)], np.int32).reshape(-1, 3)
        assert np.array_equal(k, O.fast9_16(img, th)), th
        fast["fast_%d" % th] = k
    yx = np.concatenate([rng.integers(-300000, 300000, (4000, 2)), [[0, 0], [0, 5], [5, 0], [0, -5], [-5, 0], [7, 7], [-7, 7], [7, -7], [-7, -7]]]).astype(np.int32)
    at = np.array([cv2.fastAtan2(float(y), float(x)) for y, x in yx], f32)
    assert all(at[i] == f32(O.fast_atan2(*yx[i])) for i in range(len(yx)))
    q = synth.random_descriptors(64, 1, 0.2); t = synth.random_descriptors(257, 2, 0.3); t[100] = t[7]; t[200] = q[5]; t[201] = q[5]
    knn = cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(q, t, k=2)
    kidx = np.array([[m.trainIdx for m in r] for r in knn], np.int32); kdist = np.array([[int(m.distance) for m in r] for r in knn], np.int32)
    oi, od = O.bf_knn2(q, t)
    assert np.array_equal(kidx, oi) and np.array_equal(kdist, od)
    np.savez_compressed(os.path.join(outdir, "primitives.npz"), image=img, blur=blur, atan_yx=yx, atan_deg=at,
                        bf_q=q, bf_t=t, bf_idx=kidx, bf_dist=kdist, **rs, **fast)
    print("primitives ok")


class PyVocabulary:
    """Second, independent restatement of DBoW2's vocabulary for the golden vectors: a literal node list with Python
    children lists, std::map semantics via dicts kept in key order (TemplatedVocabulary.h loadFromTextFile, transform
    :1127-1200 / :1218-1259, BowVector.cpp addWeight / normalize, FeatureVector.cpp addFeature)."""

    def __init__(self, parent, is_leaf, desc, weight, L):
        n = len(parent)
        self.children = [[] for _ in range(n)]
        for i in range(1, n):
            self.children[int(parent[i])].append(i)
        self.bits = np.unpackbits(np.asarray(desc, np.uint8), axis=1)
        self.weight = [float(w) for w in weight]
        self.word = [-1] * n
        words = 0
        for i in range(1, n):
            if is_leaf[i]:
                self.word[i] = words; words += 1
        self.L = L

    def descend(self, d, levelsup):
        bits = np.unpackbits(np.asarray(d, np.uint8))
        nid_level = self.L - levelsup
        nid = 0 if nid_level <= 0 else None
        final, level = 0, 0
        while True:
            level += 1
            nodes = self.children[final]
            final = nodes[0]
            best = int((bits != self.bits[final]).sum())
            for c in nodes[1:]:
                dd = int((bits != self.bits[c]).sum())
                if dd < best:
                    best, final = dd, c
            if level == nid_level:
                nid = final
            if not self.children[final]:
                break
        if nid is None:
            nid = final            # the reference leaves *nid uninitialised here; canonical choice: the leaf
        return self.word[final], self.weight[final], nid

    def transform(self, descs, levelsup):
        bow, fv = {}, {}
        for i, d in enumerate(descs):
            w, wt, nid = self.descend(d, levelsup)
            if wt > 0:
                if w in bow:
                    bow[w] += wt
                else:
                    bow[w] = wt
                fv.setdefault(nid, []).append(i)
        norm = 0.0
        for k in sorted(bow):
            norm += abs(bow[k])
        if norm > 0.0:
            for k in bow:
                bow[k] /= norm
        return bow, fv


def py_search_by_bow(mode, k1, d1, valid1, fv1, k2, d2, valid2, fv2, nnratio, check_ori):
    """Literal restatement of ORBmatcher::SearchByBoW (ORBmatcher.cc:269-471 mono branch / :819-959) on dict FeatureVectors."""
    TH_LOW, HL = 50, 30
    bits1 = np.unpackbits(d1, axis=1); bits2 = np.unpackbits(d2, axis=1)
    m12 = [-1] * len(k1); matched2 = [False] * len(k2)
    rot = [[] for _ in range(HL)]
    n = 0
    keys1, keys2 = sorted(fv1), sorted(fv2)
    a = b = 0
    while a < len(keys1) and b < len(keys2):
        if keys1[a] == keys2[b]:
            for i1 in fv1[keys1[a]]:
                if not valid1[i1]:
                    continue
                b1, bi, b2 = 256, -1, 256
                for i2 in fv2[keys2[b]]:
                    if matched2[i2]:
                        continue
                    if mode == 1 and not valid2[i2]:
                        continue
                    dist = int((bits1[i1] != bits2[i2]).sum())
                    if dist < b1:
                        b2, b1, bi = b1, dist, i2
                    elif dist < b2:
                        b2 = dist
                ok = b1 <= TH_LOW if mode == 0 else b1 < TH_LOW
                if ok and np.float32(b1) < np.float32(nnratio) * np.float32(b2):
                    m12[i1] = bi; matched2[bi] = True
                    if check_ori:
                        r = np.float32(k1["angle"][i1]) - np.float32(k2["angle"][bi])
                        if r < 0:
                            r = np.float32(r + np.float32(360.0))
                        x = np.float32(r * np.float32(np.float32(1.0) / np.float32(HL)))
                        bn = int(np.floor(x + np.float32(0.5))) if x >= 0 else int(np.ceil(x - np.float32(0.5)))     # C round()
                        if bn == HL:
                            bn = 0
                        rot[bn].append(i1)
                    n += 1
            a += 1; b += 1
        elif keys1[a] < keys2[b]:
            while a < len(keys1) and keys1[a] < keys2[b]:
                a += 1
        else:
            while b < len(keys2) and keys2[b] < keys1[a]:
                b += 1
    if check_ori:
        sizes = [len(r) for r in rot]
        mx = [0, 0, 0]; ind = [-1, -1, -1]
        for i, s_ in enumerate(sizes):                  # ComputeThreeMaxima, ORBmatcher.cc:2312-2353
            if s_ > mx[0]:
                mx = [s_, mx[0], mx[1]]; ind = [i, ind[0], ind[1]]
            elif s_ > mx[1]:
                mx = [mx[0], s_, mx[1]]; ind = [ind[0], i, ind[1]]
            elif s_ > mx[2]:
                mx[2] = s_; ind[2] = i
        if mx[1] < np.float32(0.1) * np.float32(mx[0]):
            ind[1] = ind[2] = -1
        elif mx[2] < np.float32(0.1) * np.float32(mx[0]):
            ind[2] = -1
        for i in range(HL):
            if i in ind:
                continue
            for i1 in rot[i]:
                m12[i1] = -1; n -= 1
    return n, np.array(m12, np.int32)


def golden_search_by_bow(outdir):
    st = synth.rects_stream(480, 360, 2, seed=31)
    e = O.Extractor(600, 1.2, 8, 20, 7)
    _, k1, d1 = e(st[0], (0, 0)); _, k2, d2 = e(st[1], (0, 0))
    vocab = synth.random_vocabulary(k=10, L=4, seed=9)
    V = O.Vocabulary(*vocab, L=4); P = PyVocabulary(*vocab, L=4)
    out = dict(k1=k1, d1=d1, k2=k2, d2=d2)
    for key, arr in zip(("parent", "leaf", "desc", "weight"), vocab):
        out["voc_" + key] = arr
    v1 = (np.arange(len(k1)) % 6 != 0).astype(np.uint8); v2 = (np.arange(len(k2)) % 5 != 0).astype(np.uint8)
    out["valid1"] = v1; out["valid2"] = v2
    for levelsup in (3, 2):
        fv1 = P.transform(d1, levelsup)[1]; fv2 = P.transform(d2, levelsup)[1]
        ofv1 = V.transform(d1, levelsup)[1]; ofv2 = V.transform(d2, levelsup)[1]
        for mode, ratio in ((0, 0.7), (1, 0.8), (0, 0.95)):
            n, m12 = py_search_by_bow(mode, k1, d1, v1, fv1, k2, d2, v2, fv2, ratio, True)
            on, om12 = O.search_by_bow(mode, k1, d1, v1, ofv1, k2, d2, v2 if mode == 1 else None, ofv2, ratio, True)
            assert n == on and np.array_equal(m12, om12), (levelsup, mode)
            n2_, m2 = py_search_by_bow(mode, k1, d1, v1, fv1, k2, d2, v2, fv2, ratio, False)
            on2, om2 = O.search_by_bow(mode, k1, d1, v1, ofv1, k2, d2, v2 if mode == 1 else None, ofv2, ratio, False)
            assert n2_ == on2 and np.array_equal(m2, om2)
            tag = "ls%d_m%d_r%d" % (levelsup, mode, int(ratio * 100))
            out[tag + "_n"] = np.int32(n); out[tag + "_m12"] = m12; out[tag + "_noori_n"] = np.int32(n2_); out[tag + "_noori_m12"] = m2
            print("search_by_bow", tag, n, n2_)
    np.savez_compressed(os.path.join(outdir, "search_by_bow.npz"), **out)


def golden_bow(outdir):
    cases = {"k10_L3": dict(k=10, L=3), "irregular_k6_L4": dict(k=6, L=4, irregular=True, shuffle=True), "wide_k20_L2": dict(k=20, L=2)}
    out = {}
    for name, kw in cases.items():
        vocab = synth.random_vocabulary(seed=7, **kw)
        q = synth.vocabulary_queries(vocab, 300, seed=8)
        L = kw["L"]
        py = PyVocabulary(*vocab, L=L)
        oc = O.Vocabulary(*vocab, L=L)
        for levelsup in (1, L - 1, L + 2):
            bow, fv = py.transform(q, levelsup)
            (bw, bv), (fn, ff) = oc.transform(q, levelsup)
            assert list(bw) == sorted(bow) and [bow[k] for k in sorted(bow)] == list(bv), name
            assert list(fn) == sorted(fv) and all(list(a) == fv[k] for a, k in zip(ff, sorted(fv))), name
            w, wt, nd = oc.transform_features(q, levelsup)
            ref = [py.descend(d, levelsup) for d in q]
            assert [r[0] for r in ref] == list(w) and [r[1] for r in ref] == list(wt) and [r[2] for r in ref] == list(nd), name
            out["%s_ls%d_word" % (name, levelsup)] = w; out["%s_ls%d_node" % (name, levelsup)] = nd
            out["%s_ls%d_bow_words" % (name, levelsup)] = bw; out["%s_ls%d_bow_values" % (name, levelsup)] = bv
        for key, arr in zip(("parent", "leaf", "desc", "weight"), vocab):
            out["%s_%s" % (name, key)] = arr
        out["%s_queries" % name] = q
        out["%s_L" % name] = np.int32(L)
    np.savez_compressed(os.path.join(outdir, "bow.npz"), **out)
    print("bow ok")


def main():
    outdir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
    os.makedirs(outdir, exist_ok=True)
    golden_primitives(outdir)
    golden_bow(outdir)
    golden_search_by_bow(outdir)
    st = synth.rects_stream(320, 240, 2, seed=11)
    k0, d0 = golden_extract("extract_320x240_nf300_f0", st[0], (300, 1.2, 8, 20, 7), (0, 0), outdir)
    k1, d1 = golden_extract("extract_320x240_nf300_f1", st[1], (300, 1.2, 8, 20, 7), (0, 0), outdir)
    golden_extract("extract_320x240_nf300_mono", st[0], (300, 1.2, 8, 20, 7), (0, 1000), outdir)
    golden_extract("extract_376x240_nf500_lap", synth.rects_frame(376, 240, 5), (500, 1.2, 8, 20, 7), (100, 250), outdir)
    golden_extract("extract_noise_256x192_nf400", synth.noise_frame(256, 192, 2), (400, 1.2, 6, 20, 7), (0, 0), outdir)
    golden_extract("extract_flat_200x160", np.full((160, 200), 90, np.uint8), (200, 1.2, 8, 20, 7), (0, 0), outdir)
    # matcher goldens are produced by the C oracle and cross-checked against cv2 where cv2 has the op
    n, m12, prev = O.search_for_initialization(k0, d0, k1, d1, (0, 320, 0, 240), np.stack([k0["x"], k0["y"]], 1), 100, 0.9, True)
    np.savez_compressed(os.path.join(outdir, "search_init_320x240.npz"), k1=k0, d1=d0, k2=k1, d2=d1,
                        nmatches=np.int32(n), matches12=m12, prev=prev)
    print("search_for_initialization nmatches", n)


if __name__ == "__main__":
    main()
