"""CPU tests of the prediction-side scorer."""
import numpy as np
import pytest

from dgdm_b200 import metrics as M


def test_predicted_metrics_thresholds_and_units():
    std, thr = np.array(M.STD["point"]), np.array(M.THRESHOLD["point"])
    lg = np.array([[(thr[0] * 1.01) / std[0], 0.0, -(thr[2] * 1.01) / std[2]],
                   [(thr[0] * 0.99) / std[0], (thr[1] * 2) / std[1], 0.0],
                   [-(thr[0] * 3) / std[0], -(thr[1] * 2) / std[1], (thr[2] * 0.5) / std[2]]])
    m = M.predicted_metrics(lg, "point")
    assert m["profile"].tolist() == [2, 1, 0]
    assert m["profile_x"].tolist() == [1, 2, 0]
    assert m["profile_y"].tolist() == [0, 1, 1]
    assert np.allclose(m["delta_theta"], lg[:, 0] * std[0] * 180 / np.pi)
    assert np.allclose(m["delta_pos"], lg[:, 1:] * std[1:])


def test_metric2objective_keys_and_selection():
    rs = np.random.RandomState(0)
    lg = rs.randn(6, 36, 3)
    for name in ("rotate", "rotate_clockwise", "shift_left", "counterclockwise_down", "clockwise_right"):
        objs = M.predicted_objectives(lg, name)
        best = M.get_best_ids_all_metrics(objs, name)
        assert "success_rate" in best
        for k, i in best.items():
            vals = [o[k] for o in objs]
            assert i == (int(np.argmin(vals)) if M._direction(k, name) == "min" else int(np.argmax(vals)))
    o = M.predicted_objectives(lg, "clockwise_up")[0]
    assert set(o) == {"success_rate", "num_clockwise_up_classes", "num_clockwise_classes", "delta_theta",
                      "num_up_classes", "delta_pos_x"}
    assert o["num_clockwise_up_classes"] == o["num_clockwise_classes"] + o["num_up_classes"]
    assert M._direction("delta_theta", "clockwise_up") == "min" and M._direction("delta_pos_x", "clockwise_up") == "min"
    assert M._direction("delta_theta", "counterclockwise_right") == "max"
    assert M._direction("delta_pos_y", "counterclockwise_right") == "max"
    with pytest.raises(ValueError, match="opt obj not supported"):
        M.metric2objective(M.predicted_metrics(lg[0]), "spin")


def test_export_designs():
    d = np.linspace(-1, 1, 28).reshape(2, 14, 1)
    e = M.export_designs(d, "point")
    assert e.shape == (2, 14, 2)
    assert np.allclose(e[0, :7, 0], np.linspace(-0.12, 0.12, 7)) and np.allclose(e[0, 7:, 0], np.linspace(-0.12, 0.12, 7))
    assert np.allclose(e[..., 1], d[..., 0] * 0.03 - 0.015)
    e3 = M.export_designs(np.zeros((3, 42, 1)), "point_3d")
    assert e3.shape == (3, 42) and np.allclose(e3, -0.05)
