"""Live cross-check of the oracle against the reference imported from /root/reference (build
container only; skipped elsewhere — the committed fixtures cover the GPU box)."""
import numpy as np
import pytest

import parity
import refharness as RH
from oracle import oracle as O
from scenarios import random_config
from zrk_modulation_b200 import compile_batch, make_noise

pytestmark = [pytest.mark.reference,
              pytest.mark.skipif(not RH.available(), reason="reference tree not present")]


@pytest.mark.parametrize("family,seed", [(f, s) for f in ("mixed", "dense", "mid", "tiny") for s in range(100, 112)])
def test_oracle_vs_live_reference(family, seed):
    cfg = random_config(seed, family)
    b = compile_batch(cfg)
    noise = make_noise(8, b.max_radars, b.n_slots, seed=seed)
    ref = RH.run_reference(cfg, noise)
    b.set_noise(noise)
    r = O.run(b, 0, b.n_ticks(0), numpy_pow=True)
    assert parity.diff_msgs(ref["msgs"], ref["ids"], r["msgs"], r["ids"]) == ""
    assert (r["fault_code"], r["fault_tick"]) == (ref["fault_code"], ref["fault_tick"]), ref["fault_where"]
    assert parity.diff_state(ref["state"], r["state"], known=ref["state_known"]) == ""


def test_ui_written_yaml_with_aliases_defaults_and_extra_keys(tmp_path):
    """a scenario file as the reference's editor writes it (UI/PolygonEditor.py:262-279,
    522-593): the missiles share their launcher's position list, so yaml.dump emits anchors and
    aliases; missiles carry an extra `position` key; optional keys are left to main.py's defaults
    (max_missiles 5, velocity 1000, explosion_radius 50, life_time 60); plus keys nobody reads.
    The file goes through the reference's own loader + object builder and through compile_batch."""
    import yaml
    lpos = [500.0, 500.0, 0.0]                       # one list object, referenced four times
    cfg = {
        "simulation": {"time_step": 200, "duration": 200 * 80},
        "air_environment": {"id": 999, "position": [0, 0, 0], "targets": [
            {"id": 40, "type": "AIR_PLANE", "position": [9000.0, 2000.0, 3000.0], "velocity": [-200.0, 0.0, 0.0], "name": "bandit"},
            {"id": 41, "type": "HELICOPTER", "position": [-4000.0, 7000.0, 1500.0], "velocity": [60.0, -70.0, 0.0]}]},
        "combat_control_point": {"id": 0, "missile_launcher_ids": [100, 101], "radar_ids": [5], "colour": "red"},
        "missile_launchers": [
            {"id": 100, "position": lpos, "max_missiles": 2, "missile_velocity": 900, "missiles": [
                {"id": 100001, "position": lpos, "velocity": 900, "explosion_radius": 120, "life_time": 30},
                {"id": 100002, "position": lpos},                                  # all defaults
                {"id": 100003, "position": lpos, "velocity": 800}]},               # beyond max_missiles: dropped
            {"id": 101, "position": [0.0, -800.0, 0.0], "missiles": [{"id": 101001}]}],   # max_missiles default 5
        "radars": [{"id": 5, "position": [0, 0, 0], "azimuth_start": 0, "elevation_start": 0, "max_distance": 30000,
                    "azimuth_range": 360, "elevation_range": 90, "azimuth_speed": 0, "elevation_speed": 0,
                    "scan_mode": "horizontal", "comment": "main radar"}],
        "ui": {"zoom": 1.5}}
    path = tmp_path / "ui_written.yaml"
    path.write_text(yaml.dump(cfg, allow_unicode=True, sort_keys=False))
    assert "&id001" in path.read_text() and "*id001" in path.read_text()      # the aliases are really there
    from zrk_modulation_b200 import load_config
    loaded = load_config(str(path))
    b = compile_batch(str(path))
    assert int(b.n_missiles[0]) == 3 and list(b.missile_id[0, :3]) == [100001, 100002, 101001]
    assert list(b.missile_v0[0, :3]) == [900.0, 1000.0, 1000.0] and list(b.missile_radius[0, :3]) == [120.0, 50.0, 50.0]
    assert list(b.missile_life[0, :3]) == [30.0, 60.0, 60.0]
    noise = make_noise(8, b.max_radars, b.n_slots, seed=3)
    ref = RH.run_reference(loaded, noise)
    b.set_noise(noise)
    r = O.run(b, 0, b.n_ticks(0), numpy_pow=True)
    assert parity.diff_msgs(ref["msgs"], ref["ids"], r["msgs"], r["ids"]) == ""
    assert (r["fault_code"], r["fault_tick"]) == (ref["fault_code"], ref["fault_tick"]), ref["fault_where"]
    assert len(ref["msgs"]) > 300
