"""Scenario compiler: reference YAML schema -> host structure-of-arrays for the C ABI.

Follows main.create_objects_from_config (main.py:35-149) key by key, including its defaults
(max_missiles=5 main.py:80; velocity=1000, explosion_radius=50, life_time=60 main.py:94-96),
the silent drop of missiles beyond max_missiles (MissileLauncher.py:43-48), the ignored
per-missile 'position' key (main.py:93) and the `objects_by_id` lookup that builds the CCP's
launcher dict (main.py:107-111).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Sequence, Union

import numpy as np

from . import _cabi

TARGET_TYPES = {"AIR_PLANE": 0, "HELICOPTER": 1, "ANOTHER": 2}  # utils.TargetType (utils.py:6-9)
SCAN_MODES = {"horizontal": 0, "vertical": 1}


def load_config(config_path: str) -> Dict[str, Any]:
    """main.load_config (main.py:30-33)."""
    import yaml
    with open(config_path, "r") as f:
        return yaml.safe_load(f)


@dataclass
class Batch:
    """B scenarios as padded SoA numpy arrays (the fields of zrk_batch_desc)."""
    dt_ms: int
    durations: np.ndarray                 # [B] simulation.duration (ms)
    n_targets: np.ndarray
    n_radars: np.ndarray
    n_launchers: np.ndarray
    n_missiles: np.ndarray
    n_ccp_launchers: np.ndarray
    airenv_id: np.ndarray
    ccp_id: np.ndarray
    target_id: np.ndarray
    target_type: np.ndarray
    target_pos: np.ndarray
    target_vel: np.ndarray
    radar_id: np.ndarray
    radar_pos: np.ndarray
    radar_par: np.ndarray
    launcher_id: np.ndarray
    launcher_pos: np.ndarray
    missile_id: np.ndarray
    missile_launcher: np.ndarray
    missile_v0: np.ndarray
    missile_v0sq: np.ndarray
    missile_radius: np.ndarray
    missile_life: np.ndarray
    ccp_launcher_id: np.ndarray
    ccp_launcher_pos: np.ndarray
    tick_limit: Optional[np.ndarray] = None
    noise: Optional[np.ndarray] = None     # [W, R, NS, 3] shared or [B, W, R, NS, 3]
    _keep: list = field(default_factory=list, repr=False)

    @property
    def n_scenarios(self) -> int:
        return int(self.n_targets.shape[0])

    @property
    def max_targets(self) -> int:
        return int(self.target_id.shape[1])

    @property
    def max_radars(self) -> int:
        return int(self.radar_id.shape[1])

    @property
    def max_launchers(self) -> int:
        return int(self.launcher_id.shape[1])

    @property
    def max_missiles(self) -> int:
        return int(self.missile_id.shape[1])

    @property
    def max_ccp_launchers(self) -> int:
        return int(self.ccp_launcher_id.shape[1])

    @property
    def n_slots(self) -> int:
        return self.max_targets + self.max_missiles

    def shard(self, s0: int, s1: int) -> "Batch":
        """Scenarios [s0, s1) as a Batch of their own (array views; a shared noise table stays shared)."""
        kw = {}
        for f, v in vars(self).items():
            if f == "_keep":
                continue
            if f == "noise":
                kw[f] = v if v is None or v.ndim == 4 else np.ascontiguousarray(v[s0:s1])
            elif isinstance(v, np.ndarray):
                kw[f] = np.ascontiguousarray(v[s0:s1])
            else:
                kw[f] = v
        return Batch(**kw)

    def dims(self):
        return (self.n_scenarios, self.n_slots, self.max_radars, self.max_missiles,
                self.max_launchers, self.max_ccp_launchers)

    def n_ticks(self, scenario: int = 0) -> int:
        """Ticks Manager.run_simulation executes: while time < duration (Manager.py:116)."""
        return int(-(-int(self.durations[scenario]) // self.dt_ms))

    def object_id(self, scenario: int, static_index: int) -> int:
        T = self.max_targets
        if static_index < T:
            return int(self.target_id[scenario, static_index])
        return int(self.missile_id[scenario, static_index - T])

    def set_noise(self, noise: Optional[np.ndarray]) -> None:
        if noise is not None:
            noise = np.ascontiguousarray(noise, dtype=np.float64)
            if noise.ndim not in (4, 5) or noise.shape[-1] != 3:
                raise ValueError("noise must be [W,R,NS,3] or [B,W,R,NS,3]")
            if noise.shape[-2] != self.n_slots or noise.shape[-3] < self.max_radars:
                raise ValueError("noise table does not cover max_radars x n_slots")
            if noise.ndim == 5 and noise.shape[0] != self.n_scenarios:
                raise ValueError("per-scenario noise needs one table per scenario")
        self.noise = noise

    def desc(self) -> _cabi.ZrkBatchDesc:
        """ctypes zrk_batch_desc whose pointers borrow this Batch's arrays."""
        d = _cabi.ZrkBatchDesc()
        d.abi_version = _cabi.ZRK_ABI_VERSION
        d.n_scenarios = self.n_scenarios
        d.max_targets, d.max_radars = self.max_targets, self.max_radars
        d.max_launchers, d.max_missiles = self.max_launchers, self.max_missiles
        d.max_ccp_launchers = self.max_ccp_launchers
        d.dt_ms = int(self.dt_ms)
        keep = []

        def p(a, dt, ct):
            a = np.ascontiguousarray(a, dtype=dt)
            keep.append(a)
            return a.ctypes.data_as(ct)

        i32 = lambda a: p(a, np.int32, _cabi._pi32)
        f64 = lambda a: p(a, np.float64, _cabi._pf64)
        for name in ("n_targets", "n_radars", "n_launchers", "n_missiles", "n_ccp_launchers",
                     "airenv_id", "ccp_id", "target_id", "target_type", "radar_id", "launcher_id",
                     "missile_id", "missile_launcher", "ccp_launcher_id"):
            setattr(d, name, i32(getattr(self, name)))
        for name in ("target_pos", "target_vel", "radar_pos", "radar_par", "launcher_pos",
                     "missile_v0", "missile_v0sq", "missile_radius", "missile_life",
                     "ccp_launcher_pos"):
            setattr(d, name, f64(getattr(self, name)))
        if self.tick_limit is not None:
            d.tick_limit = i32(self.tick_limit)
        if self.noise is not None:
            nz = self.noise
            d.noise = f64(nz)
            d.noise_period, d.noise_radars, d.noise_objects = nz.shape[-4], nz.shape[-3], nz.shape[-2]
            d.noise_scenario_stride = int(np.prod(nz.shape[1:])) if nz.ndim == 5 else 0
        self._keep = keep
        return d


def _flat_scenario(cfg: Dict[str, Any]) -> Dict[str, Any]:
    """One config dict -> python lists, walking the keys exactly as main.py does."""
    sim = cfg["simulation"]
    env = cfg["air_environment"]
    out: Dict[str, Any] = {"dt": sim["time_step"], "duration": sim["duration"], "airenv_id": env["id"]}
    by_id: Dict[Any, Any] = {env["id"]: [float(x) for x in env["position"]]}  # main.py:53

    radars = []
    for rc in cfg.get("radars", []):  # main.py:56-71
        mode = rc["scan_mode"]
        radars.append(dict(
            id=rc["id"], pos=[float(x) for x in rc["position"]],
            par=[float(rc["azimuth_start"]), float(rc["elevation_start"]), float(rc["max_distance"]),
                 float(rc["azimuth_range"]), float(rc["elevation_range"]), float(rc["azimuth_speed"]),
                 float(rc["elevation_speed"]), float(SCAN_MODES.get(mode, 2))]))
        by_id[rc["id"]] = radars[-1]["pos"]
    launchers, missiles = [], []
    for li, lc in enumerate(cfg.get("missile_launchers", [])):  # main.py:74-101
        lpos = [float(x) for x in lc["position"]]
        max_missiles = lc.get("max_missiles", 5)
        accepted = 0
        for mc in lc.get("missiles", []):
            v0 = mc.get("velocity", 1000)
            rec = dict(id=mc["id"], launcher=li, v0=float(v0), v0sq=float(v0 ** 2),
                       radius=float(mc.get("explosion_radius", 50)), life=float(mc.get("life_time", 60)))
            if accepted < max_missiles:  # MissileLauncher.add_missile (MissileLauncher.py:43-48)
                missiles.append(rec)
                accepted += 1
        launchers.append(dict(id=lc["id"], pos=lpos))
        by_id[lc["id"]] = lpos
    ccp = cfg.get("combat_control_point", {})
    ccp_launchers: List[Dict[str, Any]] = []
    if ccp:  # main.py:104-125
        seen = {}
        for ml_id in ccp.get("missile_launcher_ids", []):
            if ml_id in by_id:
                seen[ml_id] = by_id[ml_id]  # dict comprehension: first position, last value
        ccp_launchers = [dict(id=k, pos=v) for k, v in seen.items()]
        out["ccp_id"] = ccp["id"]
    else:
        out["ccp_id"] = _cabi.ZRK_NONE
    targets = []
    for tc in env.get("targets", []):  # main.py:128-147
        if tc["type"] not in TARGET_TYPES:
            raise AttributeError(tc["type"])  # getattr(TargetType, ...) main.py:129
        targets.append(dict(id=tc["id"], type=TARGET_TYPES[tc["type"]],
                            pos=[float(x) for x in tc["position"]],
                            vel=[float(x) for x in tc["velocity"]]))
    ids = [t["id"] for t in targets] + [m["id"] for m in missiles]
    if len(set(ids)) != len(ids):
        raise ValueError("unsupported scenario: air-object ids (targets + missiles) must be unique")
    lids = [l["id"] for l in launchers]
    if len(set(lids)) != len(lids):
        raise ValueError("unsupported scenario: missile launcher ids must be unique")
    out.update(radars=radars, launchers=launchers, missiles=missiles, ccp_launchers=ccp_launchers,
               targets=targets)
    return out


def compile_batch(configs: Union[Dict[str, Any], str, Sequence[Union[Dict[str, Any], str]]],
                  noise: Optional[np.ndarray] = None) -> Batch:
    """Compile one config (dict or YAML path) or a list of them into a Batch.

    All scenarios of a batch share `simulation.time_step`; counts may be ragged."""
    if isinstance(configs, (dict, str)):
        configs = [configs]
    flat = [_flat_scenario(load_config(c) if isinstance(c, str) else c) for c in configs]
    if not flat:
        raise ValueError("empty batch")
    dts = {f["dt"] for f in flat}
    if len(dts) != 1:
        raise ValueError("all scenarios of a batch must share simulation.time_step")
    B = len(flat)
    mx = lambda k: max(1, max(len(f[k]) for f in flat))
    T, R, L, M, Cc = mx("targets"), mx("radars"), mx("launchers"), mx("missiles"), mx("ccp_launchers")
    i32 = lambda *s: np.zeros(s, dtype=np.int32)
    f64 = lambda *s: np.zeros(s, dtype=np.float64)
    b = Batch(
        dt_ms=int(flat[0]["dt"]), durations=np.array([f["duration"] for f in flat], dtype=np.int64),
        n_targets=i32(B), n_radars=i32(B), n_launchers=i32(B), n_missiles=i32(B),
        n_ccp_launchers=i32(B), airenv_id=i32(B), ccp_id=i32(B),
        target_id=i32(B, T), target_type=i32(B, T), target_pos=f64(B, T, 3), target_vel=f64(B, T, 3),
        radar_id=i32(B, R), radar_pos=f64(B, R, 3), radar_par=f64(B, R, 8),
        launcher_id=i32(B, L), launcher_pos=f64(B, L, 3),
        missile_id=i32(B, M), missile_launcher=i32(B, M), missile_v0=f64(B, M), missile_v0sq=f64(B, M),
        missile_radius=f64(B, M), missile_life=f64(B, M),
        ccp_launcher_id=i32(B, Cc), ccp_launcher_pos=f64(B, Cc, 3))
    for s, f in enumerate(flat):
        b.n_targets[s], b.n_radars[s] = len(f["targets"]), len(f["radars"])
        b.n_launchers[s], b.n_missiles[s] = len(f["launchers"]), len(f["missiles"])
        b.n_ccp_launchers[s] = len(f["ccp_launchers"])
        b.airenv_id[s], b.ccp_id[s] = f["airenv_id"], f["ccp_id"]
        for i, t in enumerate(f["targets"]):
            b.target_id[s, i], b.target_type[s, i] = t["id"], t["type"]
            b.target_pos[s, i], b.target_vel[s, i] = t["pos"], t["vel"]
        for i, r in enumerate(f["radars"]):
            b.radar_id[s, i], b.radar_pos[s, i], b.radar_par[s, i] = r["id"], r["pos"], r["par"]
        for i, l in enumerate(f["launchers"]):
            b.launcher_id[s, i], b.launcher_pos[s, i] = l["id"], l["pos"]
        for i, m in enumerate(f["missiles"]):
            b.missile_id[s, i], b.missile_launcher[s, i] = m["id"], m["launcher"]
            b.missile_v0[s, i], b.missile_v0sq[s, i] = m["v0"], m["v0sq"]
            b.missile_radius[s, i], b.missile_life[s, i] = m["radius"], m["life"]
        for i, c in enumerate(f["ccp_launchers"]):
            b.ccp_launcher_id[s, i], b.ccp_launcher_pos[s, i] = c["id"], c["pos"]
    if len({int(d) for d in b.durations}) > 1:  # ragged durations: every scenario stops at its own end_time
        b.tick_limit = np.array([b.n_ticks(s) for s in range(B)], dtype=np.int32)
    b.set_noise(noise)
    return b


def make_noise(period: int, n_radars: int, n_slots: int, seed: int = 12345, sigma: float = 5.0) -> np.ndarray:
    """Shared radar-noise table of SURVEY §8d: fp64 N(0, sigma^2), shape [W, R, NS, 3]."""
    rng = np.random.default_rng(seed)
    return rng.normal(0.0, sigma, size=(period, max(1, n_radars), n_slots, 3))


def synthetic_batch(n_scenarios: int, n_targets: int, n_radars: int = 1, n_launchers: int = 0,
                    missiles_per_launcher: int = 0, seed0: int = 0, dt_ms: int = 200,
                    sweep_second_radar: bool = True, noise_period: int = 64,
                    box_km: float = 20.0) -> Batch:
    """The synthetic scenario family of SURVEY §8d, generated directly as arrays.

    Scenario s is drawn from numpy.random.default_rng(seed0 + s): positions uniform in
    x,y in [-box,box] km, z in [0.5,8] km; velocities vx,vy in [-250,250], vz in [-10,10] m/s;
    target ids 1000+i.  Radar r sits at (100 r, 0, 0) with max_distance 30 km, elevation_range 90;
    radar 0 covers 360 deg, radar 1 (if any) sweeps a 30 deg sector.  Launcher m at (500 m, 500, 0)
    with missiles velocity 1000, explosion_radius 150, life_time 60."""
    B, T, R, L = n_scenarios, n_targets, n_radars, n_launchers
    M = L * missiles_per_launcher
    pos = np.empty((B, T, 3))
    vel = np.empty((B, T, 3))
    for s in range(B):
        rng = np.random.default_rng(seed0 + s)
        pos[s, :, 0:2] = rng.uniform(-box_km * 1e3, box_km * 1e3, size=(T, 2))
        pos[s, :, 2] = rng.uniform(500.0, 8000.0, size=T)
        vel[s, :, 0:2] = rng.uniform(-250.0, 250.0, size=(T, 2))
        vel[s, :, 2] = rng.uniform(-10.0, 10.0, size=T)
    Rp, Lp, Mp = max(1, R), max(1, L), max(1, M)
    radar_par = np.zeros((B, Rp, 8))
    radar_pos = np.zeros((B, Rp, 3))
    for r in range(R):
        radar_pos[:, r, 0] = 100.0 * r
        if r == 1 and sweep_second_radar:
            radar_par[:, r] = [0.0, 0.0, 30000.0, 30.0, 90.0, 30.0, 0.0, 0.0]
        else:
            radar_par[:, r] = [0.0, 0.0, 30000.0, 360.0, 90.0, 0.0, 0.0, 0.0]
    launcher_pos = np.zeros((B, Lp, 3))
    for l in range(L):
        launcher_pos[:, l] = [500.0 * l, 500.0, 0.0]
    full = lambda v, *s: np.full(s, v)
    b = Batch(
        dt_ms=dt_ms, durations=np.zeros(B, dtype=np.int64),
        n_targets=full(T, B).astype(np.int32), n_radars=full(R, B).astype(np.int32),
        n_launchers=full(L, B).astype(np.int32), n_missiles=full(M, B).astype(np.int32),
        n_ccp_launchers=full(L, B).astype(np.int32),
        airenv_id=full(999, B).astype(np.int32), ccp_id=np.zeros(B, dtype=np.int32),
        target_id=np.tile(np.arange(1000, 1000 + T, dtype=np.int32), (B, 1)),
        target_type=np.zeros((B, T), dtype=np.int32), target_pos=pos, target_vel=vel,
        radar_id=np.tile(np.arange(5, 5 + Rp, dtype=np.int32), (B, 1)), radar_pos=radar_pos,
        radar_par=radar_par,
        launcher_id=np.tile(np.arange(100, 100 + Lp, dtype=np.int32), (B, 1)), launcher_pos=launcher_pos,
        missile_id=np.tile(np.arange(5000, 5000 + Mp, dtype=np.int32), (B, 1)),
        missile_launcher=np.tile((np.arange(Mp) // max(1, missiles_per_launcher)).astype(np.int32), (B, 1)),
        missile_v0=full(1000.0, B, Mp), missile_v0sq=full(1e6, B, Mp),
        missile_radius=full(150.0, B, Mp), missile_life=full(60.0, B, Mp),
        ccp_launcher_id=np.tile(np.arange(100, 100 + Lp, dtype=np.int32), (B, 1)),
        ccp_launcher_pos=launcher_pos.copy())
    if noise_period:
        b.set_noise(make_noise(noise_period, R, b.n_slots))
    return b


def perturbed_batch(config: Union[Dict[str, Any], str], n_scenarios: int, pos_sigma: float = 100.0,
                    vel_sigma: float = 5.0, seed: int = 0, noise: Optional[np.ndarray] = None) -> Batch:
    """Monte-Carlo family around one scenario: `n_scenarios` copies of `config` whose target
    start positions / velocities are perturbed by N(0, pos_sigma^2) / N(0, vel_sigma^2);
    scenario 0 is the unperturbed original."""
    base = compile_batch(config, noise=noise)
    rep = lambda a: np.repeat(a, n_scenarios, axis=0)
    b = Batch(**{f: (rep(v) if isinstance(v, np.ndarray) and f != "noise" else v)
                 for f, v in vars(base).items() if f not in ("_keep",)})
    rng = np.random.default_rng(seed)
    T = int(base.n_targets[0])
    b.target_pos[1:, :T] += rng.normal(0.0, pos_sigma, size=(n_scenarios - 1, T, 3))
    b.target_vel[1:, :T] += rng.normal(0.0, vel_sigma, size=(n_scenarios - 1, T, 3))
    return b


def batch_to_config(b: Batch, s: int) -> Dict[str, Any]:
    """Inverse of compile_batch for one scenario: a dict in the reference's YAML schema."""
    T = int(b.n_targets[s])
    names = {v: k for k, v in TARGET_TYPES.items()}
    modes = {0: "horizontal", 1: "vertical", 2: "other"}
    cfg: Dict[str, Any] = {
        "simulation": {"time_step": int(b.dt_ms), "duration": int(b.durations[s])},
        "air_environment": {"id": int(b.airenv_id[s]), "position": [0.0, 0.0, 0.0], "targets": [
            {"id": int(b.target_id[s, i]), "type": names[int(b.target_type[s, i])],
             "position": [float(x) for x in b.target_pos[s, i]],
             "velocity": [float(x) for x in b.target_vel[s, i]]} for i in range(T)]},
        "radars": [], "missile_launchers": []}
    for r in range(int(b.n_radars[s])):
        p = b.radar_par[s, r]
        cfg["radars"].append({
            "id": int(b.radar_id[s, r]), "position": [float(x) for x in b.radar_pos[s, r]],
            "azimuth_start": float(p[0]), "elevation_start": float(p[1]), "max_distance": float(p[2]),
            "azimuth_range": float(p[3]), "elevation_range": float(p[4]), "azimuth_speed": float(p[5]),
            "elevation_speed": float(p[6]), "scan_mode": modes[int(p[7])]})
    for l in range(int(b.n_launchers[s])):
        ms = [m for m in range(int(b.n_missiles[s])) if int(b.missile_launcher[s, m]) == l]
        cfg["missile_launchers"].append({
            "id": int(b.launcher_id[s, l]), "position": [float(x) for x in b.launcher_pos[s, l]],
            "max_missiles": max(5, len(ms)),
            "missiles": [{"id": int(b.missile_id[s, m]), "velocity": float(b.missile_v0[s, m]),
                          "explosion_radius": float(b.missile_radius[s, m]),
                          "life_time": float(b.missile_life[s, m])} for m in ms]})
    if int(b.ccp_id[s]) != _cabi.ZRK_NONE:
        cfg["combat_control_point"] = {
            "id": int(b.ccp_id[s]),
            "missile_launcher_ids": [int(x) for x in b.ccp_launcher_id[s, :int(b.n_ccp_launchers[s])]],
            "radar_ids": [int(x) for x in b.radar_id[s, :int(b.n_radars[s])]]}
    return cfg
