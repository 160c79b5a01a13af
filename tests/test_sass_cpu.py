"""CPU suite, part 6: static properties of the built sm_100a library that the measured performance rests on
(read with cuobjdump, no GPU needed): resources of the row kernel, and the instruction mix and size of the
diagonal-band column loops (7 LOP3 + 1 IADD3 + 1 SHF per window word, loop bodies small enough
that all live widths fit the instruction caches together)."""
import collections
import os
import re
import shutil
import subprocess

import pytest

import util
from isocon_b200 import _binding

ROW = "_ZN6isocon13nn_row_kernelENS_9GraphArgsEii"
pytestmark = pytest.mark.skipif(shutil.which("cuobjdump") is None or not os.path.exists(_binding.LIB_PATH),
                                reason="needs cuobjdump and the built library")


def _loops(sass):
    ins = []
    pat = re.compile(r"^\s+/\*([0-9a-f]{4,})\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_.]+)\s*(.*?);")
    for line in sass.splitlines():
        m = pat.match(line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2), m.group(3)))
    index = {a: i for i, (a, _, _) in enumerate(ins)}
    out = []
    for i, (a, op, rest) in enumerate(ins):
        if op.startswith("BRA"):
            m = re.search(r"0x([0-9a-f]+)", rest)
            if m and int(m.group(1), 16) <= a and int(m.group(1), 16) in index:
                body = ins[index[int(m.group(1), 16)]:i + 1]
                out.append(collections.Counter(o.split(".")[0] for _, o, _ in body))
    return out


def test_row_kernel_resources_keep_two_blocks_per_sm():
    txt = subprocess.check_output(["cuobjdump", "-res-usage", _binding.LIB_PATH], stderr=subprocess.DEVNULL).decode()
    m = re.search(r"Function %s:\s*\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)" % ROW, txt)
    assert m, txt[:400]
    reg, stack, shared, local = map(int, m.groups())
    assert reg <= 128, "more than 128 registers: one block of 256 threads per SM instead of two"
    assert local == 0 and stack <= 256, (stack, local)       # the DP state must live in registers


def test_column_loops_of_every_width():
    sass = subprocess.check_output(["cuobjdump", "-sass", "-fun", ROW, _binding.LIB_PATH], stderr=subprocess.DEVNULL).decode()
    # the innermost column loops of diag_segment<W>: straight-line blocks of U columns ending in one backward
    # branch, no POPC / votes / global loads inside
    found = {}
    for mix in _loops(sass):
        n = sum(mix.values())
        if mix["BRA"] != 1 or mix["POPC"] or mix["LDG"] or mix["VOTE"] or not mix["LDS"] or mix["LEA"] > 2:
            continue
        lds = mix["LDS"]
        for W in range(1, 15):
            U = 8 if W <= 2 else 4 if W <= 4 else 2 if W <= 6 else 1          # DiagUnroll<W>
            if lds == W * U and mix["LOP3"] in range(7 * W * U, 7 * W * U + 2 * U + 1):
                found[W] = (U, n, mix)
    assert sorted(found) == list(range(1, 15)), sorted(found)
    for W, (U, n, mix) in found.items():
        # per column: 7 LOP3 per word (+1 symbol mask), 1 SHF per word (+ bottom bit, + symbol shift), the adds of
        # the carry chain; whole body well below the 6 KB L0 instruction cache (16 B per instruction)
        assert mix["LOP3"] <= (7 * W + 1) * U + 1, (W, mix)
        assert mix["SHF"] <= (W + 2) * U + 2, (W, mix)
        assert mix["IADD3"] + mix["VIADD"] <= W * U + 3, (W, mix)
        assert n * 16 <= 3072, (W, n)
        assert n <= (10 * W + 6) * U + 8, (W, n)        # 9 ALU + 1 LDS per word, ~5 per column, loop control
