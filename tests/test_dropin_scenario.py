"""Class-level parity of the drop-in: EVERY public method of ORBextractor / ORBmatcher / ORBVocabulary, driven through the
reference's own Frame / KeyFrame / MapPoint code by tests/scenario/scenario.cc, must produce the same dump as the reference's own
ORBextractor.cc + ORBmatcher.cc + DBoW2 (scenario_ref = objects of oracle/_ref):
  * CPU suite: dropin/*.cc over the oracle-backed ABI stand-in -> checks the host logic of the class layer without a GPU;
  * GPU suite: dropin/*.cc over liborbx_b200.so               -> the product path, on the device.
The binaries are built where /root/reference exists (tests/scenario/Makefile, called by __graft_entry__.build()) and travel."""
import os
import subprocess

import numpy as np
import pytest

from multi_orbslam3_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "tests", "scenario", "_build")
REFROOT = "/root/reference/src/orb_slam3_ros/orb_slam3"


def binaries():
    if os.path.exists(os.path.join(REFROOT, "src", "ORBmatcher.cc")):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle", "ref")])
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tests", "scenario")])
    return all(os.path.exists(os.path.join(BUILD, b)) for b in ("scenario_ref", "scenario_dropin_cpu", "scenario_dropin_gpu"))


def make_inputs(tmp_path, W, H, nframes, seed):
    stream = synth.rects_stream(W, H, nframes, seed=seed)
    L, R = synth.stereo_pair(W, H, seed=seed + 1, disparity=14)
    np.concatenate([stream.reshape(nframes, -1), L.reshape(1, -1), R.reshape(1, -1)]).tofile(tmp_path / "frames.raw")
    vocab = synth.random_vocabulary(k=6, L=5, seed=21)       # levelsup = 4 -> FeatureVector nodes at level 1: six of them
    with open(tmp_path / "voc.txt", "w") as fh:         # the text format of ORBvoc.txt
        fh.write("6 5 0 0\n")
        rows = ["%d %d %s %r" % (vocab[0][i], vocab[1][i], " ".join(str(int(b)) for b in vocab[2][i]), float(vocab[3][i]))
                for i in range(1, len(vocab[0]))]
        fh.write("\n".join(rows))                       # no trailing newline: the reference's loader reads until eof


def run(binary, tmp_path, W, H, nframes, tag):
    out = tmp_path / ("%s.txt" % tag)
    env = dict(os.environ)
    p = subprocess.run([os.path.join(BUILD, binary), str(tmp_path / "frames.raw"), str(W), str(H), str(nframes), str(tmp_path / "voc.txt"), str(out)],
                       env=env, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    return out.read_text().splitlines(), p.stderr


def compare(ref, got):
    assert len(ref) == len(got), (len(ref), len(got))
    headline = ""
    for a, b in zip(ref, got):
        if not a.startswith(" "):
            headline = a.split(":")[0][:80]
        if a != b:
            fa, fb = a.split(), b.split()
            bad = [i for i, (x, y) in enumerate(zip(fa, fb)) if x != y]
            raise AssertionError("%s / %s: %d of %d fields differ, first at %s: %s vs %s" %
                                 (headline, fa[0], len(bad) + abs(len(fa) - len(fb)), len(fa), bad[:5], [fa[i] for i in bad[:5]], [fb[i] for i in bad[:5]]))


# every method must actually match something, or the comparison proves nothing
EXPECT_NONZERO = ("SearchForInitialization n=", "SearchByProjection(Cur,Last)[0] n=", "SearchByProjection(Cur,Last)[1] n=",
                  "SearchByProjection(Cur,Last)[2] n=", "SearchByProjection(F,MapPoints) th=3", "SearchByProjection(reloc) n=",
                  "SearchByBoW(KF,F) n=", "SearchByBoW(KF,KF) n=", "SearchByProjection(Sim3) n=", "SearchByProjection(Sim3,KFs) n=",
                  "SearchForTriangulation[0] n=", "SearchForTriangulation[2] n=", "SearchBySim3 n=", "Fuse(Sim3) n=", "Fuse n=",
                  "SearchByProjection(CurRig,Last)[0] n=", "SearchByProjection(CurRig,Last)[1] n=", "SearchByProjection(CurRig,Last)[2] n=",
                  "SearchByProjection(Rig,MapPoints) th=1", "SearchByProjection(Rig,MapPoints) th=3", "Fuse(RigKF,left) n=",
                  "SearchByBoW(KF,Rig) n=", "SearchByBoW(RigKF,Rig) n=")


def check_coverage(lines):
    for key in EXPECT_NONZERO:
        hit = [l for l in lines if l.startswith(key)]
        assert hit, key
        n = int(hit[0].rsplit("n=", 1)[1].split()[0])
        assert n > 0, hit[0]
    # the brute-force matcher of ComputeStereoFishEyeMatches: three train sets (full, one row, empty), ratio test passes somewhere
    bf = [l for l in lines if l.startswith("BFmatcher.knnMatch")]
    assert len(bf) == 3 and int(bf[0].rsplit("ratio-accepted=", 1)[1]) > 0 and " train=1 " in bf[1] and " train=0 " in bf[2], bf


@pytest.mark.skipif(not binaries(), reason="tests/scenario/_build was not prebuilt and /root/reference is absent")
def test_dropin_host_logic_equals_reference(tmp_path):
    """CPU: the drop-in classes (over the oracle-backed ABI stand-in) == the reference's own classes, every method, bit for bit."""
    W, H, NF = 752, 480, 4
    make_inputs(tmp_path, W, H, NF, seed=5)
    ref, _ = run("scenario_ref", tmp_path, W, H, NF, "ref")
    got, _ = run("scenario_dropin_cpu", tmp_path, W, H, NF, "dropin_cpu")
    check_coverage(ref)
    compare(ref, got)


@pytest.mark.gpu
@pytest.mark.skipif(not binaries(), reason="tests/scenario/_build was not prebuilt")
@pytest.mark.parametrize("W,H,seed", [(752, 480, 5), (640, 480, 9)], ids=["euroc", "tum"])
def test_dropin_on_gpu_equals_reference(tmp_path, W, H, seed):
    """GPU: the drop-in classes over liborbx_b200.so == the reference's own classes, every method, bit for bit; the device-side
    ORBmatcher::ComputeStereoMatches equals Frame::ComputeStereoMatches."""
    NF = 4
    make_inputs(tmp_path, W, H, NF, seed=seed)
    ref, _ = run("scenario_ref", tmp_path, W, H, NF, "ref")
    got, err = run("scenario_dropin_gpu", tmp_path, W, H, NF, "dropin_gpu")
    check_coverage(ref)
    compare(ref, got)
    assert "device ComputeStereoMatches identical" in err, err[-500:]
