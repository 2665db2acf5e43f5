"""CPU tests of the agglomeration row (C.dendrogram, PhyBin.hs:358): the host restatement's merge list,
dendrogram.txt's text, and the sequential model of the device kernel's protocol against the host restatement."""
import numpy as np
import pytest

from phybin_b200 import host
from tests.agglo_model import agglo_model

LINK = ["single", "complete", "UPGMA"]


def _packed(full):
    m = full.shape[0]
    return np.concatenate([full[i, i + 1:] for i in range(m - 1)]).astype(np.uint16) if m > 1 else np.zeros(0, np.uint16)


def _random_full(rng, m, hi):
    a = rng.integers(0, hi + 1, size=(m, m))
    full = np.triu(a, 1)
    return (full + full.T).astype(np.uint16)


@pytest.mark.parametrize("linkage", LINK)
def test_host_merges_match_labels_and_are_monotone(linkage):
    rng = np.random.default_rng(7)
    for m in (2, 3, 9, 40):
        full = _random_full(rng, m, 6)
        sub = _packed(full)
        labels, ma, mb, md = host.agglomerate(sub, m, linkage)
        assert len(ma) == m - 1 and np.all(ma < mb) and np.all(labels == 0)
        assert np.all(np.diff(md) >= 0) or linkage == "UPGMA" and np.all(np.diff(md) >= -1e-9)  # reducible linkages
        for thresh in (0.0, 1.0, 2.5):
            l2, a2, b2, d2 = host.agglomerate(sub, m, linkage, thresh)
            k = len(a2)
            assert np.array_equal(a2, ma[:k]) and np.array_equal(b2, mb[:k]) and (k == m - 1 or md[k] > thresh)
            assert np.array_equal(l2, host.cluster_component(sub, m, linkage, thresh))


@pytest.mark.parametrize("linkage", [0, 1, 2])
@pytest.mark.parametrize("G", [1, 2, 3, 7, 16])
def test_kernel_protocol_model_equals_host(linkage, G):
    rng = np.random.default_rng(100 * linkage + G)
    for m, hi in ((2, 1), (3, 1), (5, 2), (17, 2), (33, 3), (60, 8), (48, 0)):
        full = _random_full(rng, m, hi)   # few distinct values: ties everywhere
        for thresh in (float("inf"), 1.0):
            _, ha, hb, hd = host.agglomerate(_packed(full), m, LINK[linkage], thresh)
            ma, mb, md = agglo_model(full, linkage, thresh, G)
            assert np.array_equal(ma, ha) and np.array_equal(mb, hb)
            assert np.array_equal(md, hd)   # same IEEE operations in the same order: bitwise equal


def test_dendrogram_text_format():
    # two leaves at distance 2 joined with a third at UPGMA distance 2.5
    txt = host.dendrogram_text(["a.tr", "b\"q", "c"], [0, 0], [1, 2], [2.0, 2.5])
    assert txt == 'Branch 2.5 (Branch 2.0 (Leaf "a.tr") (Leaf "b\\"q")) (Leaf "c")'
    assert host.dendrogram_text(["x"], [], [], []) == 'Leaf "x"'
    # show :: Double
    for v, s in ((0.0, "0.0"), (1.0, "1.0"), (12.0, "12.0"), (1.75, "1.75"), (1 / 3, "0.3333333333333333"), (0.05, "5.0e-2"),
                 (0.1, "0.1"), (1e7, "1.0e7"), (9999999.0, "9999999.0"), (123456.789, "123456.789"), (2 / 3, "0.6666666666666666"),
                 (1.1e-3, "1.1e-3")):
        assert host.dendrogram_text(["p", "q"], [0], [1], [v]) == 'Branch %s (Leaf "p") (Leaf "q")' % s
    with pytest.raises(ValueError):
        host.dendrogram_text(["p", "q", "r"], [0], [1], [1.0])   # not a full dendrogram


def test_deep_chain_dendrogram_text_is_iterative():
    m = 20000
    ma = np.zeros(m - 1, np.uint32)
    mb = np.arange(1, m, dtype=np.uint32)
    txt = host.dendrogram_text(["t%d" % i for i in range(m)], ma, mb, np.arange(m - 1, dtype=np.float64))
    assert txt.count("Leaf") == m and txt.count("Branch") == m - 1


def _haskell_show_double(x: float) -> str:
    """Haskell's `show :: Double -> String` (Numeric.showFloat / floatToDigits base 10) from Python's shortest repr digits."""
    from decimal import Decimal
    if x == 0:
        return "0.0"
    sign, digits, exp = Decimal(repr(x)).as_tuple()
    ds = "".join(map(str, digits)).rstrip("0") or "0"
    e = len(digits) + exp          # x = 0.d1d2... * 10^e
    if 0.1 <= x < 1e7:
        if e <= 0:
            return "0." + "0" * (-e) + ds
        whole = (ds + "0" * e)[:e]
        frac = ds[e:] or "0"
        return whole + "." + frac
    return ds[0] + "." + (ds[1:] or "0") + "e" + str(e - 1)


def test_show_double_matches_haskell_rule_on_random_values():
    rng = np.random.default_rng(11)
    vals = np.concatenate([rng.integers(0, 1000, 200) / rng.integers(1, 400, 200),       # UPGMA-like averages
                           rng.random(200) * 10.0 ** rng.integers(-6, 9, 200), [0.1, 0.09999999999999999, 1e7, 9999999.999999998]])
    for v in vals:
        v = float(v)
        want = 'Branch %s (Leaf "p") (Leaf "q")' % _haskell_show_double(v)
        assert host.dendrogram_text(["p", "q"], [0], [1], [v]) == want, v


@pytest.mark.parametrize("linkage", LINK)
def test_slice_of_the_full_dendrogram_equals_the_cut_agglomeration(linkage):
    """sliceDendro on the full merge list (PhyBin.hs:415-422) == stopping at the first distance > thresh."""
    rng = np.random.default_rng(23)
    for m in (1, 2, 7, 40, 90):
        full = _random_full(rng, m, 5)
        sub = _packed(full)
        _, ma, mb, md = host.agglomerate(sub, m, linkage)
        for thresh in (-1.0, 0.0, 1.0, 2.5, 4.0, 100.0):
            cut, _, _, _ = host.agglomerate(sub, m, linkage, thresh)
            assert np.array_equal(host.slice_dendrogram(m, ma, mb, md, thresh), cut), (m, thresh)
