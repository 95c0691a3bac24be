"""ctypes front-end for the CPU oracle (oracle/twoarmy_oracle.c).

TEST INFRASTRUCTURE ONLY -- see the header of twoarmy_oracle.c.  Imported by
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs, never by the product package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_SO = _HERE / "_build" / "libtwoarmy_oracle.so"

# mirrors `ora_env` (C layout, natural alignment)
ENV_DTYPE = np.dtype(
    [
        ("version", np.int32),
        ("grid", np.uint8, (289,)),
        ("ax", np.int32),
        ("ay", np.int32),
        ("step_count", np.int32),
        ("step_move", np.int32),
        ("risk_count", np.int32),
        ("pone", np.uint8),
        ("patrol", np.uint8),
        ("up1", np.uint8),
        ("right2", np.uint8),
        ("upd_h", np.uint8),
        ("upd_l", np.uint8),
        ("first_room2", np.uint8),
        ("error", np.uint8),
        ("mid", np.int8, (3, 2)),
        ("o1", np.int8, (3, 2)),
        ("o2", np.int8, (4, 2)),
        ("t", np.uint32),
    ],
    align=True,
)

RNG_REPLAY, RNG_PHILOX, RNG_MT = 0, 1, 2
REWARD_LUT = np.array([-0.01, -0.1, -0.9, 0.2, 0.9], dtype=np.float64)


def build(force: bool = False) -> Path:
    """Compile the oracle with the committed Makefile (gcc only, no reference sources)."""
    src = _HERE / "twoarmy_oracle.c"
    if force or not _SO.exists() or _SO.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE)], check=True, capture_output=True)
    return _SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(str(_SO))
        L.ora_env_size.restype = C.c_size_t
        assert L.ora_env_size() == ENV_DTYPE.itemsize, (L.ora_env_size(), ENV_DTYPE.itemsize)
        L.ora_bench_rollout.restype = C.c_int64
        L.ora_np_choice.restype = C.c_int
        L.mt_next.restype = C.c_uint32
        L.ora_step_mt.restype = C.c_int
        L.ora_reward_value.restype = C.c_double
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class OracleBatch:
    """N independent oracle envs stepped in lock-step (the vector-env contract of
    DESIGN.md: terminal obs returned, then reset() on done when autoreset)."""

    def __init__(self, version: int, n: int, view: int = 17, seed: int = 0, env_id0: int = 0):
        assert version in (4, 6)
        self.version, self.n, self.view = version, int(n), int(view)
        self.seed, self.env_id0 = int(seed), int(env_id0)
        self.envs = np.zeros(self.n, dtype=ENV_DTYPE)
        lib().ora_batch_init(_p(self.envs), C.c_int64(self.n), C.c_int(version))

    def reset(self, mask=None) -> np.ndarray:
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        lib().ora_batch_reset(_p(self.envs), C.c_int64(self.n), _p(m))
        return self.obs()

    def obs(self) -> np.ndarray:
        out = np.empty((self.n, self.view, self.view, 3), dtype=np.uint8)
        lib().ora_batch_obs(_p(self.envs), C.c_int64(self.n), C.c_int(self.view), _p(out))
        return out

    def matrix(self) -> np.ndarray:
        out = np.empty((self.n, 289), dtype=np.float64)
        lib().ora_batch_matrix(_p(self.envs), C.c_int64(self.n), _p(out))
        return out

    def step(self, actions, draws=None, autoreset=True):
        """draws: uint8 [n,8] replay record (0xFF = site not executed) or None -> Philox."""
        a = np.ascontiguousarray(actions, dtype=np.int32)
        n, V = self.n, self.view
        obs = np.empty((n, V, V, 3), dtype=np.uint8)
        rew = np.empty(n, dtype=np.float32)
        ridx = np.empty(n, dtype=np.uint8)
        term = np.empty(n, dtype=np.uint8)
        trunc = np.empty(n, dtype=np.uint8)
        cons = np.empty(n, dtype=np.uint8)
        vals = np.empty((n, 8), dtype=np.uint8)
        d = None if draws is None else np.ascontiguousarray(draws, dtype=np.uint8)
        lib().ora_batch_step(
            _p(self.envs), C.c_int64(n), _p(a), C.c_int(RNG_REPLAY if d is not None else RNG_PHILOX),
            _p(d), C.c_uint64(self.seed), C.c_uint64(self.env_id0), C.c_int(V), C.c_int(int(autoreset)),
            _p(obs), _p(rew), _p(ridx), _p(term), _p(trunc), _p(cons), _p(vals),
        )
        return dict(obs=obs, reward=rew, reward_idx=ridx, terminated=term, truncated=trunc,
                    consumed=cons, values=vals)


class OracleMT:
    """One oracle env fed by the numpy-legacy MT19937 stream (np.random.seed(int) then
    np.random.choice at the reference's call sites) -- reproduces SURVEY.md section-4."""

    def __init__(self, version: int, view: int, np_seed: int):
        self.env = np.zeros(1, dtype=ENV_DTYPE)
        self.view = view
        self.mt = (C.c_uint32 * 625)()
        lib().mt_seed(self.mt, C.c_uint32(np_seed & 0xFFFFFFFF))
        lib().ora_batch_init(_p(self.env), C.c_int64(1), C.c_int(version))

    def obs(self):
        out = np.empty((1, self.view, self.view, 3), dtype=np.uint8)
        lib().ora_batch_obs(_p(self.env), C.c_int64(1), C.c_int(self.view), _p(out))
        return out[0]

    def reset(self):
        lib().ora_batch_reset(_p(self.env), C.c_int64(1), None)
        return self.obs()

    def step(self, action: int):
        V = self.view
        obs = np.empty((V, V, 3), dtype=np.uint8)
        te, tr = C.c_int(0), C.c_int(0)
        cons = C.c_uint8(0)
        vals = np.empty(8, dtype=np.uint8)
        ri = lib().ora_step_mt(_p(self.env), C.c_int(int(action)), self.mt, C.c_int(V), _p(obs),
                               C.byref(te), C.byref(tr), C.byref(cons), _p(vals))
        return obs, float(REWARD_LUT[ri]), bool(te.value), bool(tr.value), cons.value, vals

    def encode_grid(self):
        out = np.empty((17, 17, 3), dtype=np.uint8)
        lib().ora_encode_grid(_p(self.env), _p(out))
        return out


def gen_obs_general(grid, ax, ay, agent_dir, view, see_through_walls):
    g = np.ascontiguousarray(grid, dtype=np.uint8)
    out = np.empty((view, view, 3), dtype=np.uint8)
    lib().ora_gen_obs_general(_p(g), C.c_int(ax), C.c_int(ay), C.c_int(agent_dir), C.c_int(view),
                              C.c_int(int(see_through_walls)), _p(out))
    return out


def philox4x32_10(ctr, key):
    c = np.asarray(ctr, dtype=np.uint32)
    k = np.asarray(key, dtype=np.uint32)
    out = np.empty(4, dtype=np.uint32)
    lib().philox4x32_10(_p(c), _p(k), _p(out))
    return out


def td_advantage(r, v, v_next, gamma):
    r = np.ascontiguousarray(r, np.float32); v = np.ascontiguousarray(v, np.float32)
    vn = np.ascontiguousarray(v_next, np.float32)
    adv = np.empty_like(r); tv = np.empty_like(r)
    lib().ora_td_advantage(_p(r), _p(v), _p(vn), C.c_float(gamma), C.c_int64(r.size), _p(adv), _p(tv))
    return adv, tv


def gae(r, v, done, gamma, lam, use_mask=True, v_next=None, last_v=None):
    """float64 GAE over [T,N]; see ora_gae."""
    r = np.ascontiguousarray(r, np.float32); v = np.ascontiguousarray(v, np.float32)
    T, N = r.shape
    d = np.ascontiguousarray(done, np.uint8)
    vn = None if v_next is None else np.ascontiguousarray(v_next, np.float32)
    lv = None if last_v is None else np.ascontiguousarray(last_v, np.float32)
    adv = np.empty((T, N), np.float64); ret = np.empty((T, N), np.float64)
    lib().ora_gae(_p(r), _p(v), _p(vn), _p(lv), _p(d), C.c_double(gamma), C.c_double(lam),
                  C.c_int(int(use_mask)), C.c_int(T), C.c_int64(N), _p(adv), _p(ret))
    return adv, ret


def max_threads() -> int:
    return int(lib().ora_max_threads())


def bench_rollout(version, n, T, view, seed=9981, threads=None):
    """Timed CPU baseline: returns (env_steps, seconds, threads_used)."""
    import time
    if threads:
        lib().ora_set_threads(C.c_int(int(threads)))
    b = OracleBatch(version, n, view, seed)
    obs = np.empty((n, view, view, 3), np.uint8)
    rew = np.empty(n, np.float32); te = np.empty(n, np.uint8); tr = np.empty(n, np.uint8)
    L = lib()
    L.ora_bench_rollout(_p(b.envs), C.c_int64(n), C.c_int(2), C.c_int(view), C.c_uint64(seed),
                        _p(obs), _p(rew), _p(te), _p(tr))  # warm-up
    t0 = time.perf_counter()
    steps = L.ora_bench_rollout(_p(b.envs), C.c_int64(n), C.c_int(T), C.c_int(view), C.c_uint64(seed),
                                _p(obs), _p(rew), _p(te), _p(tr))
    return int(steps), time.perf_counter() - t0


# ---- hindsight relabelling (soa/env_buffer.py:101-143) ------------------------------------------
def her_select(p4, newgoal_size_in=4, choice=None):
    """The selection half of Buffer_gridworld.her_func for ONE episode (env_buffer.py:108-117):
    p4 [L,2] = buffer['p'][:,4,0:2] of the episode's records.  Returns (indices, episode_idxs):
    `indices` = first-occurrence record index of every distinct position, in np.unique's (sorted
    lexicographic) order; `episode_idxs` = np.random.choice(indices, size=min(4, len), replace=False)
    drawn from the GLOBAL legacy numpy stream like the reference (or from `choice` if given)."""
    p4 = np.asarray(p4)
    _, indices, _ = np.unique(p4[:, 0:2], return_index=True, return_counts=True, axis=0)
    k = min(newgoal_size_in, indices.size)
    if p4.shape[0] <= 0:
        return indices, np.zeros(0, np.int64)
    chooser = choice or (lambda a, size: np.random.choice(a, size=size, replace=False))
    return indices, np.asarray(chooser(indices, k), dtype=np.int64)


def her_func(p4, r, counter_start, capacity=2048, newgoal_size_in=4, choice=None):
    """Buffer_gridworld.her_func (env_buffer.py:101-143) for an episode whose L records sit at
    buffer positions counter_start .. counter_start+L-1 (no wrap inside the episode, as the
    reference's slice requires).  Returns dict(src, g, r, d, counter, full): the appended records in
    order (src = index of the copied episode record, g = relabelled goal, r / d after the
    overrides of lines 126-127), the new buffer counter and whether the ring wrapped."""
    p4 = np.asarray(p4, np.float32)
    r = np.asarray(r, np.float32)
    L = p4.shape[0]
    end = counter_start + L - 1           # epo_counter_end = self.counter - 1
    src, g, rr, dd = [], [], [], []
    full = False
    if L > 0:
        _, idxs = her_select(p4, newgoal_size_in, choice)
        for index in idxs:
            index = int(index)
            if index > 0 and index < capacity:
                n = index + 1
                goal = p4[index, 0:2]
                rn = r[:n].copy(); rn[index] = np.float32(0.9)
                dn = np.zeros(n, np.float32); dn[index] = 1
                src += list(range(n)); g += [goal] * n; rr += list(rn); dd += list(dn)
                if end + 1 + n <= capacity:
                    end = end + 1 + index
                else:
                    end = end + 1 + n - capacity - 1
                    full = True
    return dict(src=np.array(src, np.int64), g=np.array(g, np.float32).reshape(-1, 2), r=np.array(rr, np.float32),
                d=np.array(dd, np.float32), counter=end + 1, full=full)


def pre_her_func(p8, r_step, newgoal_size_in=4, choice=None):
    """Buffer_gridworld.pre_her_func (env_buffer.py:145-210) for ONE episode of the PPO + predictor loop, stated
    over the episode's STEPS.  The reference's 9-frame record j (stored from the 5th step on, then four closing
    pads, train_ppo_predictor.py:140-171) holds step j + 4 in its newest slot; p8 [J,2] = pre_buffer['p'][:,8,0:2]
    of the episode's J records, r_step [L] the per-step rewards (L = J).  Returns the appended records as
    (step, g, r0) -- `step` = the step whose 5-frame record sits in frames 0..4 (what PPO_Predictor.update reads,
    PPO_Predictor.py:124-163), g the relabelled goal, r0 = r[:,0] after the override of line 170 has been
    shifted into slot 0 by the four pad records (:173-196)."""
    p8 = np.asarray(p8, np.float32)
    J = p8.shape[0]
    steps, g, r0 = [], [], []
    if J > 0:
        _, idxs = her_select(p8, newgoal_size_in, choice)
        for index in idxs:
            index = int(index)
            if index > 0:
                n = index + 1 + 4                              # records 0..index, then the 4 pads
                rn = np.asarray(r_step[:n], np.float32).copy()
                rn[n - 1] = np.float32(0.9)
                steps += list(range(n)); g += [p8[index, 0:2]] * n; r0 += list(rn)
    return dict(step=np.array(steps, np.int64), g=np.array(g, np.float32).reshape(-1, 2), r0=np.array(r0, np.float32))


def her_plan(pos_y, pos_x, done, choose, first=0):
    """first = 0: her_func; first = 4: pre_her_func (candidates are the records from the episode's 5th step on,
    indices stay relative to the episode start).
    The vectorised form the CUDA kernel computes: pos_y/pos_x/done [T,N]; every episode segment
    that ENDS inside the window (segments start at t=0 or after a done) is relabelled like her_func;
    choose(indices, k, t_end, env) -> the k chosen record indices.  Returns plan uint16 [T,N,4]:
    0xFFFF = record not in relabel slot c, else goal y*32+x | 0x8000 on the prefix's last record."""
    T, N = done.shape
    plan = np.full((T, N, 4), 0xFFFF, np.uint16)
    for e in range(N):
        t0 = 0
        for t1 in range(T):
            if not done[t1, e]:
                continue
            L = t1 - t0 + 1
            p4 = np.stack([pos_y[t0:t1 + 1, e], pos_x[t0:t1 + 1, e]], 1)
            if L <= first:
                t0 = t1 + 1
                continue
            _, indices, _ = np.unique(p4[first:], return_index=True, return_counts=True, axis=0)
            indices = indices + first
            k = min(4, indices.size)
            chosen = choose(indices, k, t1, e)
            for c, index in enumerate(chosen):
                index = int(index)
                if index > first:
                    goal = int(p4[index, 0]) * 32 + int(p4[index, 1])
                    plan[t0:t0 + index + 1, e, c] = goal
                    plan[t0 + index, e, c] = goal | 0x8000
            t0 = t1 + 1
    return plan
