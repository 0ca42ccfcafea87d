"""CPU tests of the oracle itself: pinned against the reference's own golden vectors
(/root/reference/tests/test_alignment.rs) and frozen against tests/golden/oracle_goldens.json."""
import numpy as np
import pytest

from conftest import CONFIG_TOML, TEST_CONFIG, random_pair, read_fasta_gz


def test_reference_vectors_faithful_and_linear(oracle, ref_vectors):
    sc = ref_vectors["scores"]
    scores = (sc["s_match"], sc["s_mismatch"], sc["g"], sc["h"])
    assert scores == TEST_CONFIG
    for case in ref_vectors["cases"]:
        for fn in (oracle.align_faithful, oracle.align_linear):
            r = fn(case["s1"], case["s2"], scores, ref_vectors["is_local"])
            if case["score"] is not None:
                assert r.score == case["score"]
            assert r.matches == case["matches"]
            assert r.mismatches == case["mismatches"]
            assert r.opening_gaps == case["opening_gaps"]
            assert r.gap_extensions == case["gap_extensions"]
            assert [list(x) for x in r.alignment] == case["alignment"]


@pytest.mark.parametrize("fixture", ["test1", "test2_short", "test3_short", "test4", "Opsin1_colorblindness_gene"])
@pytest.mark.parametrize("is_local", [False, True])
def test_fixture_goldens(oracle, goldens, fixture, is_local):
    g = next(p for p in goldens["pairs"] if p["fixture"] == fixture and p["is_local"] == is_local)
    s = read_fasta_gz(fixture)
    fn = oracle.align_faithful if g["m"] * g["n"] < 2e7 else oracle.align_linear
    r = fn(s[0][1], s[1][1], CONFIG_TOML, is_local)
    assert r.score == g["score"] and list(r.start) == g["start"] and list(r.end) == g["end"]
    assert (r.matches, r.mismatches, r.gap_extensions, r.opening_gaps) == (
        g["matches"], g["mismatches"], g["gap_extensions"], g["opening_gaps"])
    assert "%016x" % oracle.hash_ops(r.ops, r.start) == g["op_hash"]
    assert list(r.first_max) == g["first_max"] and r.lcs_at_first_max == g["lcs_at_first_max"]


@pytest.mark.parametrize("is_local", [False, True])
def test_brca2_linear_golden(oracle, goldens, is_local):
    g = next(p for p in goldens["pairs"] if p["fixture"] == "Human-Mouse-BRCA2-cds" and p["is_local"] == is_local)
    s = read_fasta_gz("Human-Mouse-BRCA2-cds")
    r = oracle.align_linear(s[0][1], s[1][1], CONFIG_TOML, is_local)
    assert r.score == g["score"] and list(r.start) == g["start"] and list(r.end) == g["end"]
    assert len(r.ops) == g["n_ops"] and "%016x" % oracle.hash_ops(r.ops, r.start) == g["op_hash"]


def test_faithful_equals_linear_random(oracle):
    rng = np.random.default_rng(7)
    for scores in [CONFIG_TOML, TEST_CONFIG, (2, -1, -1, 0), (5, -4, -3, -10), (1, 0, -1, -1), (3, 1, -2, -2)]:
        for _ in range(60):
            m, n = int(rng.integers(0, 40)), int(rng.integers(0, 40))
            a, b = random_pair(rng, m, n, similar=bool(rng.integers(0, 2)))
            for is_local in (False, True):
                f = oracle.align_faithful(a, b, scores, is_local)
                l = oracle.align_linear(a, b, scores, is_local)
                assert (f.score, f.start, f.end) == (l.score, l.start, l.end)
                assert np.array_equal(f.ops, l.ops) and np.array_equal(f.ops_i, l.ops_i) and np.array_equal(f.ops_j, l.ops_j)
                assert (f.matches, f.mismatches, f.gap_extensions, f.opening_gaps) == (
                    l.matches, l.mismatches, l.gap_extensions, l.opening_gaps)
                assert f.first_max == l.first_max and f.lcs_at_first_max == l.lcs_at_first_max
                sc, si, sj = oracle.score_linear(a, b, scores, is_local)
                assert sc == f.score and (si, sj) == f.start


def test_edge_cases(oracle):
    # empty vs empty emits one Match at (0,0): algo.rs:351-369 with is_match(None, None) == true
    r = oracle.align_faithful("", "", CONFIG_TOML, False)
    assert r.score == 0 and r.alignment == [("Match", 0, 0)]
    r = oracle.align_faithful("", "ACG", CONFIG_TOML, False)
    assert r.score == -5 - 3 and [x[0] for x in r.alignment] == ["OpenInsert", "Insert", "Insert"]
    r = oracle.align_faithful("ACG", "", CONFIG_TOML, False)
    assert r.score == -8 and [x[0] for x in r.alignment] == ["OpenDelete", "Delete", "Delete"]
    r = oracle.align_faithful("", "ACG", CONFIG_TOML, True)
    assert r.score == 0 and r.start == (0, 3) and len(r.ops) == 0
    r = oracle.align_faithful("AAAA", "TTTT", CONFIG_TOML, True)   # all zero: last cell wins, run-on walk
    assert r.score == 0 and r.start == (4, 4)


def test_blocked_and_batch(oracle):
    rng = np.random.default_rng(3)
    a, b = random_pair(rng, 3000, 2800)
    sc, _, _ = oracle.score_linear(a, b, CONFIG_TOML, False)
    assert oracle.nw_score_blocked(a, b, CONFIG_TOML, n_threads=4, blk=256) == sc
    assert oracle.nw_score_blocked(a, b, CONFIG_TOML, n_threads=1, blk=4096) == sc
    pairs = [random_pair(rng, int(rng.integers(1, 160)), int(rng.integers(1, 160))) for _ in range(300)]
    blob = np.concatenate([np.concatenate(p) for p in pairs])
    off1, len1, off2, len2, pos = [], [], [], [], 0
    for x, y in pairs:
        off1.append(pos); len1.append(len(x)); pos += len(x)
        off2.append(pos); len2.append(len(y)); pos += len(y)
    for is_local in (False, True):
        got = oracle.score_batch(blob, off1, len1, off2, len2, CONFIG_TOML, is_local, n_threads=4)
        exp = [oracle.score_linear(x, y, CONFIG_TOML, is_local)[0] for x, y in pairs]
        assert list(got) == exp


def test_corona_golden_sample(oracle, goldens):
    # one full 30 kb pair through the linear oracle (~10 s); the other 44 are frozen in the golden file
    order = goldens["corona_order"]
    g = next(c for c in goldens["corona"] if c["pair"] == [6, 7])
    s1 = read_fasta_gz(order[6])[0][1]
    s2 = read_fasta_gz(order[7])[0][1]
    r = oracle.align_linear(s1, s2, CONFIG_TOML, False)
    assert r.score == g["score"] and len(r.ops) == g["n_ops"]
    assert "%016x" % oracle.hash_ops(r.ops, r.start) == g["op_hash"]


def test_config4_fixture_is_the_oracle(oracle):
    """tests/golden/config4_scores.npz (SURVEY 8d config-4 parity sets) is what the oracle computes: a slice of both sets
    with the batch scorer, a few pairs with the faithful 48-byte-cell variant"""
    import os
    import numpy as np
    from conftest import CONFIG_TOML, GOLDEN
    from genomics_rs_b200 import workloads as wl
    gold = np.load(os.path.join(GOLDEN, "config4_scores.npz"))
    assert gold["parity"].size == wl.CONFIG4_PARITY_PAIRS and gold["strided"].size == 10_000_000 // wl.CONFIG4_STRIDE
    blob, off1, len1, off2, len2 = wl.reads150(0, 4000, parity_set=True)
    assert np.array_equal(oracle.score_batch(blob, off1, len1, off2, len2, CONFIG_TOML, True, n_threads=4), gold["parity"][:4000])
    for q in (0, 1, 17, 3999):
        a, b = blob[int(off1[q]):int(off1[q]) + 150], blob[int(off2[q]):int(off2[q]) + 150]
        assert oracle.align_faithful(a, b, CONFIG_TOML, True).score == gold["parity"][q]
    idx = np.arange(0, 10_000_000, wl.CONFIG4_STRIDE, dtype=np.uint64)[-3000:]
    blob, off1, len1, off2, len2 = wl.reads150_pairs(idx)
    assert np.array_equal(oracle.score_batch(blob, off1, len1, off2, len2, CONFIG_TOML, True, n_threads=4), gold["strided"][-3000:])
