"""Host-side logic of the round-2 fast paths that needs no GPU: the lazily materialised meta dict of rasterization()
and the capacity prediction of the intersection stage."""
from gsplat_b200 import ops
from gsplat_b200.rendering import _Lazy, _Meta


def test_meta_resolves_lazy_values_on_every_read_path():
    calls = []

    def make():
        calls.append(1)
        return "ids"

    m = _Meta({"a": 1, "isect_ids": _Lazy(make)})
    assert "isect_ids" in m and len(m) == 2 and calls == []  # membership / size do not materialise
    assert m["a"] == 1 and calls == []
    assert m["isect_ids"] == "ids" and m["isect_ids"] == "ids" and calls == [1]  # computed once, then stored
    for read in (lambda d: dict(d), lambda d: {**d}, lambda d: dict(d.items()), lambda d: d.copy(), lambda d: {k: d[k] for k in d}):
        m = _Meta({"a": 1, "isect_ids": _Lazy(lambda: "ids")})
        assert read(m) == {"a": 1, "isect_ids": "ids"}
    m = _Meta({"isect_ids": _Lazy(lambda: "ids")})
    assert m.get("isect_ids") == "ids" and m.get("missing", 7) == 7
    m = _Meta({"isect_ids": _Lazy(lambda: "ids")})
    assert list(m.values()) == ["ids"]
    m = _Meta({"isect_ids": _Lazy(lambda: "ids")})
    assert m.pop("isect_ids") == "ids" and m.pop("isect_ids", None) is None


def _predictor():
    p = object.__new__(ops._IsectPredictor)  # the constructor allocates pinned memory / a stream: not needed here
    p.hist, p.hits, p.misses = [], 0, 0
    return p


def test_isect_capacity_prediction():
    p = _predictor()
    assert p.capacities(1000) is None  # nothing seen yet: the first call of a shape reads the totals first
    p.observe(100000, 5000, 40)
    cap_vis, cap_isects, max_tiles = p.capacities(10**6)
    assert cap_vis >= 5000 and cap_isects >= 100000 and max_tiles == 40
    assert cap_vis <= 5000 * 1.04 + 512 and cap_isects <= 100000 * 1.04 + 2048  # a few percent of padding, not more
    assert p.capacities(5100)[0] == 5100  # never more rows than exist
    # the capacities follow the maximum of the last eight calls (views of a trainer differ) and forget older ones
    p.observe(300000, 9000, 200)
    assert p.capacities(10**6)[1] >= 300000 and p.capacities(10**6)[2] == 200
    for _ in range(8):
        p.observe(50000, 2000, 10)
    cap_vis, cap_isects, max_tiles = p.capacities(10**6)
    assert cap_isects < 60000 and cap_vis < 3000 and max_tiles == 10
    # an empty view must not produce zero-sized speculative launches
    q = _predictor()
    q.observe(0, 0, 0)
    assert q.capacities(1000) is None
