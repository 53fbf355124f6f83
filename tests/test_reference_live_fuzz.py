"""Live differential test of the host functions: where the reference checkout is present (this container; never the GPU
box, and nothing GPU-marked depends on it) the reference's OWN myolo_utils -- executed in a subprocess with its
third-party imports stubbed, exactly like the golden-vector generator does -- and the package run through the same 60
seeded random cases (tests/golden/fuzz_cases.py): decode_one_yolo_output, NMB, bbox_iou, bbox_iou_2, extract_bboxes and
the BatchGenerator encoding must agree to the last bit.  Skipped when /root/reference does not exist; the committed golden
vectors (tests/test_reference_golden.py) are what travels."""
import os
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/myolo/myolo_utils.py"


@pytest.mark.skipif(not os.path.exists(REF), reason="the reference checkout is only present in the build container")
def test_host_functions_agree_with_the_reference_on_random_cases(tmp_path):
    n = 60
    out = str(tmp_path / "ref_fuzz.npz")
    env = dict(os.environ, PYTHONPATH="")                    # the subprocess must not see this package's `myolo`
    subprocess.check_call([sys.executable, os.path.join(HERE, "golden", "make_reference_fixtures.py"), "--fuzz", out, str(n)],
                          env=env, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    ref = np.load(out)
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import fuzz_cases
    from myolo import myolo_utils as U
    assert U.__file__.startswith(os.path.dirname(HERE))       # the package, not the reference
    mine = fuzz_cases.run(U, n)
    assert set(mine) == set(ref.files) and len(mine) > 5 * n
    for k in sorted(mine):
        a, b = np.asarray(mine[k]), ref[k]
        assert a.shape == b.shape and a.dtype == b.dtype, (k, a.shape, b.shape, a.dtype, b.dtype)
        assert np.array_equal(a, b, equal_nan=a.dtype.kind == "f"), k
    # the cases are not vacuous
    assert sum(mine["dec%d" % k].shape[0] for k in range(n)) > n
    assert sum(len(mine["nmb%d" % k]) for k in range(n)) < sum(len(mine["iou_%d" % k]) for k in range(n))     # something was suppressed
    assert any((mine["eb%d" % k][3] == 0).all() for k in range(n))                                               # the empty channel
