"""Random scenario generator (reference YAML schema) shared by the golden-fixture script and
the parity tests.  Deterministic in `seed`; families stress different parts of the tick."""
from __future__ import annotations

import numpy as np


def random_config(seed: int, family: str = "mixed"):
    rng = np.random.default_rng(seed)
    u = rng.uniform
    dense = family == "dense"
    box = 1500.0 if dense else (6000.0 if family == "mid" else 20000.0)
    n_targets = int(rng.integers(0, 3)) if family == "tiny" else int(rng.integers(2, 14))
    n_radars = int(rng.integers(0, 4)) if family in ("mixed", "tiny") else int(rng.integers(1, 3))
    n_launchers = int(rng.integers(0, 4))
    dt = int(rng.choice([100, 200, 250, 500]))
    ticks = int(rng.integers(40, 140))
    cfg = {"simulation": {"time_step": dt, "duration": dt * ticks},
           "air_environment": {"id": 999, "position": [0.0, 0.0, 0.0], "targets": []},
           "radars": [], "missile_launchers": []}
    for i in range(n_targets):
        stationary = rng.random() < 0.04
        vel = [0.0, 0.0, 0.0] if stationary else [float(u(-300, 300)), float(u(-300, 300)), float(u(-15, 15))]
        z = float(u(-200, 200)) if rng.random() < 0.1 else float(u(300, 8000))
        cfg["air_environment"]["targets"].append({
            "id": 1000 + i, "type": str(rng.choice(["AIR_PLANE", "HELICOPTER", "ANOTHER"])),
            "position": [float(u(-box, box)), float(u(-box, box)), z], "velocity": vel})
    for r in range(n_radars):
        kind = rng.integers(0, 4)
        par = dict(azimuth_start=0.0, elevation_start=0.0, max_distance=float(u(8000, 60000)),
                   azimuth_range=360.0, elevation_range=float(rng.choice([90.0, 180.0])),
                   azimuth_speed=0.0, elevation_speed=0.0, scan_mode="horizontal")
        if kind == 1:   # sweeping sector
            w = float(rng.choice([30.0, 45.0, 90.0, 180.0]))
            par.update(azimuth_range=w, azimuth_speed=float(rng.choice([w, w / 2, 10.0])),
                       azimuth_start=float(rng.choice([0.0, 15.0, 90.0])))
        elif kind == 2:  # vertical mode with elevation steps
            par.update(scan_mode="vertical", elevation_range=float(rng.choice([30.0, 60.0])),
                       elevation_speed=float(rng.choice([0.0, 15.0, 30.0])),
                       azimuth_range=float(rng.choice([120.0, 360.0])), azimuth_speed=float(rng.choice([0.0, 60.0])))
        elif kind == 3:  # horizontal with elevation stepping + unknown mode sometimes
            par.update(azimuth_range=float(rng.choice([120.0, 200.0])), azimuth_speed=120.0,
                       elevation_range=45.0, elevation_speed=float(rng.choice([0.0, 20.0, 45.0])),
                       elevation_start=float(rng.choice([0.0, 10.0])),
                       scan_mode=str(rng.choice(["horizontal", "horizontal", "spiral"])))
        cfg["radars"].append({"id": 5 + r, "position": [float(u(-2000, 2000)), float(u(-2000, 2000)), float(u(0, 50))], **par})
    mid = 0
    for l in range(n_launchers):
        n_m = int(rng.integers(0, 6))
        lc = {"id": 100 + l, "position": [float(u(-3000, 3000)), float(u(-3000, 3000)), 0.0], "missiles": []}
        if rng.random() < 0.5:
            lc["max_missiles"] = int(rng.integers(1, 6))
        for _ in range(n_m):
            mc = {"id": 5000 + mid}
            mid += 1
            if rng.random() < 0.8:
                mc["velocity"] = [1000, 800.0, 300, 1234.5, 250][int(rng.integers(0, 5))]
            if rng.random() < 0.8:
                mc["explosion_radius"] = [150, 50.0, 300, 20][int(rng.integers(0, 4))]
            if rng.random() < 0.8:
                mc["life_time"] = [60, 35, 6, 3, 12.5][int(rng.integers(0, 5))]
            lc["missiles"].append(mc)
        cfg["missile_launchers"].append(lc)
    if rng.random() < 0.93:
        ids = [100 + l for l in range(n_launchers)]
        if ids and rng.random() < 0.2:
            ids = ids[::-1]
        if rng.random() < 0.1:
            ids = ids + [424242]          # unknown id: dropped (main.py:110)
        if n_radars and rng.random() < 0.1:
            ids = ids + [5]               # a radar id in the launcher list: asked, never answers
        cfg["combat_control_point"] = {"id": 0, "missile_launcher_ids": ids,
                                       "radar_ids": [5 + r for r in range(n_radars)]}
    return cfg


def special_config(kind: str):
    """Hand-made scenarios for the corners of the association scan and the visibility filter."""
    cfg = random_config(7, "mid")
    cfg["simulation"] = {"time_step": 200, "duration": 200 * 60}
    t = cfg["air_environment"]["targets"]
    if kind == "twins":          # identical tracks: exact distance ties on every tick
        t[1]["position"], t[1]["velocity"] = list(t[0]["position"]), list(t[0]["velocity"])
        t[3]["position"], t[3]["velocity"] = list(t[2]["position"]), list(t[2]["velocity"])
    elif kind == "near_twins":   # tracks a few ulps apart: near-ties, the exact path must decide
        t[1]["position"] = [t[0]["position"][0] * (1 + 2e-16), t[0]["position"][1], t[0]["position"][2]]
        t[1]["velocity"] = list(t[0]["velocity"])
        t[3]["position"] = [t[2]["position"][0], t[2]["position"][1] + 1e-9, t[2]["position"][2]]
        t[3]["velocity"] = list(t[2]["velocity"])
    elif kind == "far_field":    # large coordinates (hundreds of km): the fp32 pre-filter near the end of its range
        for i, tt in enumerate(t):
            tt["position"] = [2.0e5 + 3000.0 * i, -1.5e5 + 1000.0 * i, 9000.0]
        for r in cfg["radars"]:
            r["max_distance"] = 1.0e6
            r.update(azimuth_start=0.0, azimuth_range=360.0, azimuth_speed=0.0, elevation_start=0.0,
                     elevation_range=180.0, elevation_speed=0.0, scan_mode="horizontal")
    elif kind == "far_close":    # near the end of the fp32 range (|x| + |y| + |z| just under 2^20) AND tracks closer than fp32 resolves there
        for i, tt in enumerate(t):
            j = i // 2
            tt["position"] = [5.2e5 + 4000.0 * j + 0.03 * (i % 2), -4.6e5 + 1500.0 * j - 0.05 * (i % 2), 9000.0 + 0.02 * (i % 2)]
            if i % 2:
                tt["velocity"] = list(t[i - 1]["velocity"])
        for r in cfg["radars"]:
            r["max_distance"] = 3.0e6
            r.update(azimuth_start=0.0, azimuth_range=360.0, azimuth_speed=0.0, elevation_start=0.0,
                     elevation_range=180.0, elevation_speed=0.0, scan_mode="horizontal")
    elif kind == "beyond_f32":   # coordinates beyond the range the fp32 pre-filter is trusted in: exact scan only
        for i, tt in enumerate(t):
            tt["position"] = [3.0e6 + 3000.0 * i, -2.5e6 + 1000.0 * i, 9000.0]
        for r in cfg["radars"]:
            r["max_distance"] = 1.0e7
            r.update(azimuth_start=0.0, azimuth_range=360.0, azimuth_speed=0.0, elevation_start=0.0,
                     elevation_range=180.0, elevation_speed=0.0, scan_mode="horizontal")
    elif kind == "on_axis":      # objects exactly on sector edges / coordinate axes of the radar
        r = cfg["radars"][0]
        r.update(position=[0.0, 0.0, 0.0], azimuth_start=0.0, azimuth_range=90.0, azimuth_speed=0.0,
                 elevation_start=0.0, elevation_range=90.0, elevation_speed=0.0, max_distance=5000.0,
                 scan_mode="horizontal")
        t[0].update(position=[1000.0, 0.0, 0.0], velocity=[10.0, 0.0, 0.0])      # az = 0, el = 0
        t[1].update(position=[0.0, 1000.0, 500.0], velocity=[0.0, 10.0, 0.0])    # az = 90
        t[2].update(position=[3000.0, 4000.0, 0.0], velocity=[0.0, 0.0, 0.0])    # distance = max_distance, stationary
        t[3].update(position=[0.0, 0.0, 1000.0], velocity=[0.0, 0.0, 5.0])       # straight overhead
        if len(t) > 4:
            t[4].update(position=[0.0, 0.0, 0.0], velocity=[1.0, 1.0, 1.0])      # starts at the radar: distance 0
    elif kind == "late_launch":  # targets enter radar range after > 100 ticks: launches with a large clock value
        cfg["simulation"] = {"time_step": 200, "duration": 200 * 260}
        cfg["radars"] = [{"id": 5, "position": [0.0, 0.0, 0.0], "azimuth_start": 0.0, "elevation_start": 0.0,
                          "max_distance": 9000.0, "azimuth_range": 360.0, "elevation_range": 90.0,
                          "azimuth_speed": 0.0, "elevation_speed": 0.0, "scan_mode": "horizontal"}]
        cfg["missile_launchers"] = [{"id": 100, "position": [300.0, 200.0, 0.0], "missiles": [
            {"id": 5000 + i, "velocity": 900, "explosion_radius": 120, "life_time": 40} for i in range(3)]}]
        cfg["combat_control_point"] = {"id": 0, "missile_launcher_ids": [100], "radar_ids": [5]}
        del t[3:]
        t[0].update(position=[16000.0, 1000.0, 2000.0], velocity=[-250.0, 0.0, 0.0])   # in range after ~28 s
        t[1].update(position=[-2000.0, 19000.0, 3000.0], velocity=[10.0, -260.0, 0.0])  # ~38 s
        t[2].update(position=[9000.0, 12000.0, 1500.0], velocity=[-150.0, -200.0, 0.0])  # ~25 s
    else:
        raise ValueError(kind)
    return cfg


SPECIAL_KINDS = ["twins", "near_twins", "far_field", "far_close", "beyond_f32", "on_axis", "late_launch"]
