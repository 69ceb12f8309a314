"""f-1 pinned to the reference: dgdm_b200.metrics against tables produced by the reference's own
``metric2objective`` (dynamics/metrics.py:67-233), ``Diffusion.get_best_ids_all_metrics`` / ``get_average_best_ids`` /
``get_best_ids`` (generator/diffusion.py:346-428) on seeded simulator-style metric dicts, all 15 non-convergence
objectives (tests/golden/make_golden.py metrics -> golden_metrics.json; candidates 2 and 5 duplicate 0 and 1, so the
first-occurrence tie rule is pinned as well)."""
import json
import os

import numpy as np
import pytest

from dgdm_b200 import metrics as M


@pytest.fixture(scope="module")
def gm(golden_dir):
    with open(os.path.join(golden_dir, "golden_metrics.json")) as f:
        g = json.load(f)
    g["metrics"] = [{k: np.asarray(v) for k, v in m.items()} for m in g["metrics"]]
    return g


def test_all_15_objectives_are_covered(gm):
    assert len(gm["objectives"]) == 15
    assert set(gm["objectives"]) == set(M._SPEC) | {"rotate"}


@pytest.mark.parametrize("name", ["rotate", "rotate_clockwise", "rotate_counterclockwise", "shift_up", "shift_down",
                                  "shift_left", "shift_right", "clockwise_up", "clockwise_down", "clockwise_left",
                                  "clockwise_right", "counterclockwise_up", "counterclockwise_down",
                                  "counterclockwise_left", "counterclockwise_right"])
def test_metric2objective_and_selection_match_the_reference(gm, name):
    ref = gm["objectives"][name]
    # with the roll-out fields present the tables are the reference's, key for key, value for value, in its key order
    mine = [M.metric2objective(m, name) for m in gm["metrics"]]
    for got, want in zip(mine, ref["per_candidate"]):
        assert list(got) == list(want), (list(got), list(want))
        for k, v in want.items():
            assert got[k] == pytest.approx(v, rel=1e-6, abs=1e-9), (name, k)
            if ref["dtypes"][k] in ("int16", "int64", "int"):
                assert isinstance(got[k], int)
    best = M.get_best_ids_all_metrics(mine, name)
    assert best == ref["best_ids_all_metrics"]
    assert list(best) == list(ref["best_ids_all_metrics"])          # same keys in the same order
    assert M.get_average_best_ids(mine, name) == ref["average_best_id"]
    # prediction-only metric dicts (no final_* fields): the same tables minus the roll-out keys, same winners
    pred = [M.metric2objective({k: v for k, v in m.items() if not k.startswith("final")}, name) for m in gm["metrics"]]
    for got, want in zip(pred, ref["per_candidate"]):
        assert list(got) == [k for k in want if not k.startswith("final")]
    best_p = M.get_best_ids_all_metrics(pred, name)
    assert best_p == {k: v for k, v in ref["best_ids_all_metrics"].items() if not k.startswith("final")}


def test_get_best_ids_blocks(gm):
    objs = [M.metric2objective(m, "rotate_clockwise") for m in gm["metrics"][:6]]
    assert M.get_best_ids(objs, 3, 2, "rotate_clockwise") == gm["get_best_ids_rotate_clockwise_3x2"]


def test_errors_match_the_reference():
    with pytest.raises(ValueError, match="opt obj not supported"):
        M.get_best_ids_all_metrics([{"success_rate": 1.0}], "spin")
    with pytest.raises(ValueError, match="opt obj not supported"):
        M.get_average_best_ids([{"success_rate": 1.0}], "spin")
