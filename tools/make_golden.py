"""Writes the golden GAF fixtures under tests/golden/example/expected/: the output of the CPU oracle (the restatement of
the reference, see oracle/oracle.hpp) on the reference's shipped example for every mode and a few flag sets. The Rust
reference cannot be built in this image, so these are ORACLE outputs — regression pins for the oracle itself
(tests/test_oracle_golden.py re-derives them) and reference vectors the GPU path is compared with
(tests/test_gpu_parity.py::test_example_against_committed_fixtures).

  python tools/make_golden.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import oracle_lib  # noqa: E402

EX = os.path.join(ROOT, "tests", "golden", "example")
CASES = {
    "m0_b50": ["-m", "0", "-b", "50"],
    "m1": ["-m", "1"],
    "m2_default": ["-m", "2"],
    "m2_b50": ["-m", "2", "-b", "50"],
    "m2_s_true_b50": ["-m", "2", "-s", "true", "-b", "50"],
    "m0_s_true_b50": ["-m", "0", "-s", "true", "-b", "50"],
    "m3": ["-m", "3"],
    "m4": ["-m", "4"],
    "m5": ["-m", "5"],
    "m6": ["-m", "6"],
    "m7": ["-m", "7"],
    "m8": ["-m", "8"],
    "m9": ["-m", "9"],
}


def main():
    out_dir = os.path.join(EX, "expected")
    os.makedirs(out_dir, exist_ok=True)
    for name, flags in CASES.items():
        rc, out, err = oracle_lib.run_cli(flags + [os.path.join(EX, "reads.fa"), os.path.join(EX, "graph.gfa")])
        assert rc == 0, (name, err)
        with open(os.path.join(out_dir, name + ".gaf"), "w") as f:
            f.write(out)
        print(name, len(out.splitlines()), "lines")


if __name__ == "__main__":
    main()
