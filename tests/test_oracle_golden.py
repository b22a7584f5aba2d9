"""The C restatement against the committed golden vectors (reference-binary outputs).  Runs without
/root/reference and without the reference binaries."""
import numpy as np
import pytest

import goldens
import support


@pytest.fixture(scope="module")
def oracle():
    return support.Oracle()


@pytest.mark.parametrize("case", goldens.overlap_cases(), ids=lambda c: c["name"])
def test_overlap_golden(oracle, case):
    for (op, flags), want in case["expect"].items():
        fn = oracle.count if op == "count" else oracle.coverage
        rc, got, _ = fn(case["queries"], case["index"], flags, qw=case["qw"], qoff=case["qoff"], ioff=case["ioff"])
        assert rc == 0 and np.array_equal(got, want), (case["name"], op, flags)


@pytest.mark.parametrize("case", goldens.scan_cases(), ids=lambda c: c["name"])
def test_scan_golden(oracle, case):
    n, got = oracle.scan_counts(case["reads"], case["bounds"], case["win_step"], case["win_size"], case["op"],
                                case["ignore_strand"], case["min_reads"], weight=case["rw"])
    assert n == len(case["expect"]["value"])
    for k in ("chrom", "strand", "win", "value"):
        assert np.array_equal(got[k], case["expect"][k]), (case["name"], k)
