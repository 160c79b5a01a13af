"""CPU suite, part 5: the explicit-pair path (isocon_b200/edlib_alignment_module.py).
The fixture tests/golden/pairs_n200.json comes from the UNMODIFIED reference module
(oracle/make_golden_pairs.py); here the oracle's arithmetic is checked against it and the mirror's
interface against the reference's.  The device side is tested in test_gpu_parity.py."""
import inspect
import json
import os

import util
from isocon_b200 import workloads
from oracle import oracle as O


def _fixture():
    with open(os.path.join(util.GOLD, "pairs_n200.json")) as fh:
        return json.load(fh)


def test_oracle_distances_equal_the_reference_fixture():
    fx = _fixture()
    S = util.load_reads(200)
    Sp, _ = workloads.round1_call(S)
    n = 0
    for q, rows in fx["edlib_align_sequences"]:
        for t, ed in rows:
            assert O.ed_myers64(Sp[q].encode(), Sp[t].encode(), -1) == ed
            n += 1
    X, C = util.two_set_split(S)
    for c, rows in fx["edlib_align_sequences_keeping_accession"]:
        for r, ed in rows:
            assert O.ed_banded_dp(C[c].encode(), X[r].encode(), ed) == ed
            assert ed == 0 or O.ed_banded_dp(C[c].encode(), X[r].encode(), ed - 1) == -1
            n += 1
    assert n == 455


def test_mirror_has_the_reference_interface():
    from isocon_b200 import edlib_alignment_module as m
    want = {   # /root/reference/modules/edlib_alignment_module.py:10, :51, :103, :107, :130
        "edlib_align_sequences": ["matches", "nr_cores"],
        "edlib_align_sequences_keeping_accession": ["matches", "nr_cores"],
        "edlib_alignment_helper": ["arguments"],
        "edlib_alignment": ["x", "y", "i", "j", "x_acc", "y_acc"],
        "edlib_traceback": ["x", "y", "mode", "task", "k"],
    }
    for name, args in want.items():
        assert list(inspect.signature(getattr(m, name)).parameters) == args, name
    ref_file = "/root/reference/modules/edlib_alignment_module.py"
    if os.path.exists(ref_file):     # authoring container: compare with the reference source itself
        import ast
        tree = ast.parse(open(ref_file).read())
        ref = {f.name: [a.arg for a in f.args.args] for f in tree.body if isinstance(f, ast.FunctionDef)}
        assert ref == want
