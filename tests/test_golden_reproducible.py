"""The committed fixtures under tests/golden/ are what oracle/gen_golden.py writes today: the generator (the second,
independent restatement over real cv2 primitives, cross-checked against the C oracle while it runs) is re-run into a
temporary directory and compared array by array."""
import glob
import importlib.util
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fixtures_regenerate_identically(tmp_path):
    pytest.importorskip("cv2")
    spec = importlib.util.spec_from_file_location("gen_golden", os.path.join(ROOT, "oracle", "gen_golden.py"))
    g = importlib.util.module_from_spec(spec); spec.loader.exec_module(g)
    from multi_orbslam3_b200 import synth
    out = str(tmp_path)
    g.golden_primitives(out)
    g.golden_bow(out)
    st = synth.rects_stream(320, 240, 2, seed=11)
    g.golden_extract("extract_320x240_nf300_f0", st[0], (300, 1.2, 8, 20, 7), (0, 0), out)
    files = sorted(glob.glob(os.path.join(out, "*.npz")))
    assert len(files) == 3
    for f in files:
        a = np.load(f); b = np.load(os.path.join(ROOT, "tests", "golden", os.path.basename(f)))
        assert set(a.files) == set(b.files), os.path.basename(f)
        for k in a.files:
            np.testing.assert_array_equal(a[k], b[k], err_msg="%s:%s" % (os.path.basename(f), k))
