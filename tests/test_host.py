"""CPU tests of the host side: the C-ABI library loads and exports what include/gxalign.h declares, the
reference-shaped host mirror (FASTA, config, Display, is_match) behaves like the reference's code,
and the product path fails loudly without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, read_fasta_gz

import genomics_rs_b200 as gx
from genomics_rs_b200 import _lib


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "gxalign.h")).read()
    declared = sorted(set(re.findall(r"\b(gx_[a-z0-9_]+)\s*\(", header)))
    assert declared, "no prototypes found"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in gxalign.h but not exported"
    assert sorted(_lib.EXPORTS) == declared
    assert b"sm_100a" in lib.gx_version()
    assert C.sizeof(_lib.GxResult) == 104 and C.sizeof(_lib.GxScores) == 16


def test_status_strings_and_score_checks():
    lib = _lib.load()
    assert lib.gx_strerror(0) == b"ok"
    ok = _lib.GxScores(1, -2, -1, -5)
    assert lib.gx_check_scores(ok, 11382, 10346) == 0
    assert lib.gx_check_scores(_lib.GxScores(1, -1, -1, 1), 10, 10) == 2      # h > 0
    assert lib.gx_check_scores(_lib.GxScores(1, -2, 0, 0), 10, 10) == 2       # g == 0
    assert lib.gx_check_scores(_lib.GxScores(1, -2, -1, -5), 1 << 29, 10) == 3
    assert lib.gx_check_scores(_lib.GxScores(1 << 20, -2, -1, -5), 4000, 4000) == 3


def test_replay_ops_matches_checked_sub_rules():
    lib = _lib.load()
    ops = np.array([0, 4, 2, 5, 3, 1, 0], np.uint8)
    oi = np.zeros(7, np.uint32); oj = np.zeros(7, np.uint32)
    assert lib.gx_replay_ops(ops.ctypes.data, 7, 3, 4, oi.ctypes.data, oj.ctypes.data) == 0
    assert oi.tolist() == [3, 2, 2, 2, 1, 0, 0] and oj.tolist() == [4, 3, 2, 1, 1, 1, 0]


@pytest.mark.skipif(_lib.load().gx_device_count() > 0, reason="GPU present")
def test_product_path_fails_loudly_without_gpu():
    with pytest.raises(_lib.GxError) as e:
        gx.align_batch([("ACGT", "ACGT")], (1, -2, -2, -5), False)
    assert e.value.status == 5
    plan = C.c_void_p()
    l1 = np.array([4], np.uint64)
    assert _lib.load().gx_plan_create(l1.ctypes.data, l1.ctypes.data, 1, _lib.GxScores(1, -2, -1, -5), 0, 1, C.byref(plan)) == 8


def test_fasta_loader(tmp_path):
    p = tmp_path / "x.fasta"
    p.write_bytes(b"ACGT\n>  first seq  \r\nAC GT \n\nacgt\n>second\nTT\n")
    sc = gx.SequenceContainer()
    sc.from_fasta(str(p))
    assert [(s.name, s.sequence) for s in sc.sequences] == [("first seq", "AC GTacgt"), ("second", "TT")]
    sc.from_fasta(str(tmp_path / "missing.fasta"))          # logs an error, adds nothing (sequence.rs:85)
    assert len(sc.sequences) == 2
    import gzip
    raw = gzip.open(os.path.join(GOLDEN, "fasta", "Human-Mouse-BRCA2-cds.fasta.gz"), "rb").read()
    q = tmp_path / "brca2.fasta"
    q.write_bytes(raw)
    sc = gx.SequenceContainer()
    sc.from_fasta(str(q))
    ref = read_fasta_gz("Human-Mouse-BRCA2-cds")
    assert [len(s.sequence) for s in sc.sequences] == [11382, 10346]
    assert [(s.name, s.sequence) for s in sc.sequences] == ref


def test_is_match_option_semantics():
    sc = gx.SequenceContainer([gx.Sequence("a", "ACGT"), gx.Sequence("b", "AGCGT")])
    assert sc.is_match(0, 0) and not sc.is_match(1, 1)
    assert sc.is_match(4, 5)            # None == None
    assert not sc.is_match(4, 4)        # None vs Some
    assert not sc.is_match(3, 5)


def test_config(tmp_path):
    p = tmp_path / "config.toml"
    p.write_text("[scores]\ns_match = 1\ns_mismatch = -2\ng = -1\nh = -5")
    cfg = gx.get_config(str(p))
    assert cfg.scores == gx.Scores(1, -2, -1, -5)
    with pytest.raises(SystemExit) as e:
        gx.get_config(str(tmp_path / "nope.toml"))
    assert e.value.code == 1
    p.write_text("[scores]\ns_match = 1\n")
    with pytest.raises(SystemExit):
        gx.get_config(str(p))


def _aligned(s1, s2, ops, score, counts):
    return gx.AlignedSequences(s1=gx.Sequence("s1", s1), s2=gx.Sequence("s2", s2), score=score, matches=counts[0],
                               mismatches=counts[1], gap_extensions=counts[2], opening_gaps=counts[3],
                               ops=np.array(ops, np.uint8), start=(len(s1), len(s2)))


def test_display_matches_reference_format():
    # SURVEY.md 8f N1 example: ACGT / AGCGT, walk order [Match,Match,Match,OpenInsert,Mismatch]
    a = _aligned("ACGT", "AGCGT", [0, 0, 0, 4, 1], -3, (3, 1, 0, 1))
    text = str(a)
    assert text == ("\n\n0-5:\n\nA-CGT\nx%|||\nAGCGT\n"
                    "\n\nAlignment Score: -3\nMatches: 3/5 (60.00%)\nMismatches: 1/5 (20.00%)\n"
                    "Gap Extensions: 0/5 (0.00%)\nOpening Gaps: 1/5 (20.00%)\nPercent Identity 60%\n")
    # chunking: a new block after 201 columns (display.rs:34-44), the last header uses s1_out.len()
    n = 450
    b = _aligned("A" * n, "A" * n, [0] * n, n, (n, 0, 0, 0))
    t = str(b)
    assert t.startswith("\n\n1-201:\n\n" + "A" * 201 + "\n" + "|" * 201 + "\n" + "A" * 201 + "\n")
    assert "\n\n202-402:\n\n" in t and "\n\n402-450:\n\n" in t
    assert t.endswith("Percent Identity 100%\n")
    c = _aligned("ACG", "ACG", [0, 0, 1], 0, (2, 1, 0, 0))
    assert "Percent Identity 66.66666666666666%\n" in str(c)


def test_fasta_dir_ingestion(tmp_path):
    """main.rs:227-239: every *.fasta file, all of its records; sorted file-name order is our contract"""
    (tmp_path / "b.fasta").write_text(">b1\nACGT\n>b2\nAC\n")
    (tmp_path / "a.fasta").write_text(">a\nGG\nTT\n")
    (tmp_path / "c.txt").write_text(">x\nAA\n")
    (tmp_path / "fasta").write_text(">y\nAA\n")          # no extension: skipped like Path::extension() == None
    sc = gx.SequenceContainer()
    sc.from_fasta_dir(str(tmp_path))
    assert [(s.name, s.sequence) for s in sc.sequences] == [("a", "GGTT"), ("b1", "ACGT"), ("b2", "AC")]


def test_alignment_table_visualiser_text(oracle):
    """display.rs:131-220 with planes from the oracle (host formatting only): the reference's second test pair"""
    import numpy as np
    from genomics_rs_b200.display import format_alignment_table, format_scores_table
    s1, s2, scores = "ACGT", "AGCGT", (1, -2, -2, -5)
    o = oracle.align_faithful(s1, s2, scores, False)
    a = gx.AlignedSequences(s1=gx.Sequence("s1", s1), s2=gx.Sequence("s2", s2), score=o.score, matches=o.matches,
                            mismatches=o.mismatches, gap_extensions=o.gap_extensions, opening_gaps=o.opening_gaps,
                            ops=o.ops, start=o.start, end=o.end)
    planes = oracle.planes(s1, s2, scores, False)
    txt = format_alignment_table(a, planes)
    lines = txt.split("\n")
    assert lines[1] == "Sequence Table (S1 columns, S2 rows):" and lines[3] == " AGCGT"
    # alignment (walk order): Match(4,5) Match(3,4) Match(2,3) OpenInsert(1,2) Mismatch(1,1)  (tests/test_alignment.rs:76-89)
    assert lines[4:8] == ["AXI...", "C..M..", "G...M.", "T....M"]
    assert lines[8] == "Delete Scores" and lines[9] == ". \t0\t1\t2\t3\t4\t5\t"
    assert lines[10].startswith("0\t0\t-inf\t-inf") and lines[11].startswith("1\t-7\t")
    assert "Insert Scores" in lines and "Sub Scores" in lines
    assert format_scores_table(np.array([[0, -9223372036854775801], [5, -3]], np.int64)) == ". \t0\t1\t\n0\t0\t-inf\t\n1\t5\t-3\t\n"
    big = gx.AlignedSequences(s1=gx.Sequence("a", "A" * 200), s2=gx.Sequence("b", "A"), score=0, matches=0, mismatches=0,
                              gap_extensions=0, opening_gaps=0, ops=np.zeros(0, np.uint8))
    assert format_alignment_table(big, planes) is None      # display.rs:139-143: too large, skipped


def test_ticket_order_respects_dependencies():
    """the persistent fill kernel is deadlock-free because every tile's dependencies -- (panel, strip-1), (panel-1, strip)
    and, for column bands, the last strip of the band to the left -- are handed out before it (gx_debug_tile_order
    is the host arithmetic plan_create uses; no GPU needed)"""
    import ctypes as C
    import numpy as np
    lib = _lib.load()
    rng = np.random.default_rng(7)
    for trial in range(30):
        bands = trial % 3 == 0
        n_pairs = int(rng.integers(1, 9))
        K = int(rng.choice([4, 8, 16]))
        len2 = rng.integers(0, 6000, size=n_pairs).astype(np.uint64)
        len1 = (np.full(n_pairs, rng.integers(1, 30000), np.uint64) if bands else rng.integers(0, 30000, size=n_pairs).astype(np.uint64))
        if bands:
            len2 = np.maximum(len2, 1)
        nt = C.c_uint64()
        assert lib.gx_debug_tile_order(len1.ctypes.data, len2.ctypes.data, n_pairs, K, int(bands), None, 0, C.byref(nt)) == 0
        out = np.zeros(3 * max(nt.value, 1), np.uint32)
        assert lib.gx_debug_tile_order(len1.ctypes.data, len2.ctypes.data, n_pairs, K, int(bands), out.ctypes.data, nt.value, C.byref(nt)) == 0
        t = out[:3 * nt.value].reshape(-1, 3)
        pos = {(int(q), int(p), int(s)): k for k, (q, p, s) in enumerate(t)}
        assert len(pos) == nt.value                                     # every tile exactly once
        S = [int(-(-int(len2[q]) // (32 * K))) if len1[q] and len2[q] else 0 for q in range(n_pairs)]
        Pn = [int(-(-int(len1[q]) // 4096)) if len1[q] and len2[q] else 0 for q in range(n_pairs)]
        assert nt.value == sum(a * b for a, b in zip(S, Pn))
        for (q, p, s), k in pos.items():
            if s > 0:
                assert pos[(q, p, s - 1)] < k
            if p > 0:
                assert pos[(q, p - 1, s)] < k
            if bands and s == 0 and q > 0:
                assert pos[(q - 1, p, S[q - 1] - 1)] < k
