"""Host-side logic of the sweep driver (mpcmmd_b200/driver.py = the reference's main_mpc.py workflow): sharding, record packing,
acceptance filter / episode ordering, on-disk schema, and the one cross-rank exchange (world-size-2 gloo).  No GPU needed."""
import os
import sys

import numpy as np
import pytest

f32 = np.float32


@pytest.fixture(scope="module")
def driver(built):
    from mpcmmd_b200 import driver
    return driver


class _Prob:
    ker_wt = 1000.0


def test_thresholds_and_cost_dispatch(driver):
    assert driver.thresholds(_Prob, "mmd_opt") == (-1999.0, -999.0)          # main_mpc.py:88-89
    assert driver.thresholds(_Prob, "mmd_random") == (-1999.0, -999.0)       # :92-93
    assert driver.thresholds(_Prob, "cvar") == (1e-5, 1e-5)                  # :96-97
    assert driver.solve_cost("saa") == "cvar" and driver.solve_cost("mmd_opt") == "mmd_opt"    # the reference's else-branch


def test_shards_partition_the_sweep(driver):
    for n, w in ((200, 1), (200, 8), (7, 4), (3, 8)):
        parts = [driver.shard(n, r, w) for r in range(w)]
        assert sorted(k for p in parts for k in p) == list(range(n))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def test_data_path_matches_reference_format(driver):
    ref = "./data/{}_noise/noise_{}/ts_{}/{}_{}_samples_{}_obs".format("beta", int(0.3 * 100), 50, "mmd_opt", 5, 4)       # main_mpc.py:130-132
    assert os.path.normpath(driver.data_path("./data", "beta", 0.3, 50, "mmd_opt", 5, 4)) == os.path.normpath(ref)


def _fake_out(episodes):
    E = len(episodes); k = np.asarray(episodes, f32)
    return dict(cost_obs=np.where(k % 3 == 0, 5.0, 0.0).astype(f32), cost_lane=(k * 0.5).astype(f32),
                cx=(k[:, None] + np.arange(11)[None]).astype(f32), cy=(-k[:, None] + np.arange(11)[None]).astype(f32))


def test_assemble_filters_and_keeps_episode_order(driver):
    eps = [5, 2, 9, 0, 3, 7]
    rec = driver.pack_records(eps, _fake_out(eps))
    arrays = driver.assemble(rec, 4, 1e-5)
    kept = [k for k in sorted(eps) if k % 3 != 0]
    assert arrays["cx"].dtype == np.float64 and arrays["cx"].shape == (len(kept), 11)
    assert arrays["cx"][:, 0].tolist() == [float(k) for k in kept]                       # serial-loop order, rejected episodes dropped
    from mpcmmd_b200 import scenes
    for row, k in enumerate(kept):
        (x, y, vx, vy, _), _ = scenes.static_scene(4, k)
        assert np.array_equal(arrays["x_obs"][row], x) and np.array_equal(arrays["y_obs"][row], y)
        assert arrays["init_state"][row].tolist() == [0.0, 1.75, 5.0, 0.0, 0.0, 0.0]
    assert set(arrays) == {"cx", "cy", "init_state", "x_obs", "y_obs", "vx_obs", "vy_obs"}   # main_mpc.py:133-135


def test_static_scene_matches_reference_generator(driver):
    """compute_obs_data of main_mpc.py:10-21 + the idx_mpc draw of :113-119, restated inline"""
    from mpcmmd_b200 import scenes
    for k in (0, 1, 17, 199):
        np.random.seed(k)
        x = np.random.choice(np.array([35, 40, 45, 50, 55, 60, 65, 70, 75]), (4,), replace=False)
        y = np.random.choice(np.array([-1.75, 1.75]), (4,))
        idx = np.random.randint(1, 10000)
        (sx, sy, svx, svy, spsi), sidx = scenes.static_scene(4, k)
        assert np.array_equal(sx, x) and np.array_equal(sy, y) and sidx == idx and not svx.any() and not svy.any()
    assert scenes.static_scene(2, 0)[1] == 6745


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(1, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "mpc-mmd_b200"))
    from mpcmmd_b200 import driver
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    mine = driver.shard(7, rank, world)
    rec = driver.pack_records(mine, _fake_out(mine))
    allrec = driver.dist_gather(rec, world, "cpu")
    if rank == 0:
        arrays = driver.assemble(allrec, 2, 1e-5)
        q.put((allrec.shape, arrays["cx"][:, 0].tolist()))
    dist.barrier(); dist.destroy_process_group()


def test_episode_sharding_gloo(driver):
    """world_size 2: each rank packs its own (ragged: 4 and 3) episodes, the gather + rank-0 assembly reproduce the serial sweep"""
    import socket
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    shape, order = q.get(timeout=120)
    [p.join(timeout=60) for p in ps]
    assert all(p.exitcode == 0 for p in ps)
    assert shape == (7, 26)
    assert order == [1.0, 2.0, 4.0, 5.0]                    # episodes 0..6 minus the rejected multiples of 3, in episode order


# ---- dynamic variant: scene generator + .npz schema --------------------------------------------------------------------------------
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dynamic_scenes.npz")


@pytest.mark.parametrize("num_obs,k", [(6, 0), (6, 1), (6, 17), (6, 199), (3, 5), (2, 42)])
def test_dynamic_scene_matches_reference_generator(driver, num_obs, k):
    """scenes.dynamic_scene vs the reference's own obs_data class run on the JAX shim (tests/golden/make_golden_scenes.py).
    Initial states and idx_mpc are integer/bit-level work (Threefry bits + stable sort): exact.  Trajectories come out of a constant
    KKT solve (float32 LU in the reference, float64 inverse here): 2e-4 absolute on positions up to 150 m (= float32 round-off)."""
    from mpcmmd_b200 import scenes
    g = np.load(GOLD)
    tag = f"o{num_obs}_k{k}_"
    (x, y, vx, vy, psi), idx, xt, yt = scenes.dynamic_scene(num_obs, k)
    assert np.array_equal(x, g[tag + "x"]) and np.array_equal(vx, g[tag + "vx"]) and np.array_equal(y, g[tag + "y"])
    assert not vy.any() and not psi.any() and idx == int(g[tag + "idx"])
    assert np.abs(xt - g[tag + "xt"]).max() < 2e-4 and np.abs(yt - g[tag + "yt"]).max() < 2e-4
    assert xt.dtype == np.float32 and xt.shape == (num_obs, 100)


def test_host_jax_rng_matches_oracle(driver):
    """mpcmmd_b200.jaxrng (product-side scene RNG) vs the oracle's restatement of the same protocol, incl. the documented known answers"""
    from mpcmmd_b200 import jaxrng
    from oracle import oracle as O
    assert jaxrng.split(jaxrng.prng_key(0), 2).tolist() == [[4146024105, 967050713], [2718843009, 1272950319]]      # JAX PRNG docs
    assert jaxrng.threefry2x32((0x13198a2e, 0x03707344), np.uint32(0x243f6a88), np.uint32(0x85a308d3)) == (0xc4923a9c, 0x483df7a0)   # Random123 KAT
    for seed in (0, 5, 123, 43 * 199 + 11 * 5 + 5):
        for n in (1, 7, 30):
            assert np.array_equal(jaxrng.bits(jaxrng.prng_key(seed), n), O.bits(O.prng_key(seed), n))
            assert np.allclose(jaxrng.normal(jaxrng.prng_key(seed), n), O.normal(O.prng_key(seed), n), rtol=0, atol=5e-7)


def test_assemble_dynamic_schema(driver):
    eps = [4, 1, 2]
    rec = driver.pack_records(eps, _fake_out(eps))
    arrays = driver.assemble(rec, 3, 1e-5, "dynamic")
    assert set(arrays) == {"cx", "cy", "init_state", "x_obs", "y_obs", "vx_obs", "vy_obs", "psi_obs", "x_obs_traj", "y_obs_traj"}     # D/main_mpc.py:150-156
    assert arrays["x_obs_traj"].shape == (3, 3, 100) and arrays["psi_obs"].shape == (3, 3)
    assert arrays["init_state"][0].tolist() == [0.0, -1.75, 5.0, 0.0, 0.0, 0.0]                                    # D/main_mpc.py:34-42
    assert arrays["cx"][:, 0].tolist() == [1.0, 2.0, 4.0]


def test_empty_shard_when_more_ranks_than_episodes(driver):
    """WORLD_SIZE > --num_configs: a rank with no episodes builds correctly shaped empty inputs / records instead of crashing before the gather"""
    from mpcmmd_b200 import scenes

    class _P:
        num_obs = 4
    assert driver.shard(3, 5, 8) == []
    b = scenes.static_batch(_P, [], "static")
    assert b["idx_mpc"].shape == (0,) and b["x_obs_traj"].shape == (0, 4, 100) and b["cov_param"].shape == (0, 8, 8)
    rec = driver.pack_records([], dict(cost_obs=np.zeros(0, f32), cost_lane=np.zeros(0, f32), cx=np.zeros((0, 11), f32), cy=np.zeros((0, 11), f32)))
    assert rec.shape == (0, 26)
    full = np.concatenate([rec, driver.pack_records([0, 1, 2], _fake_out([0, 1, 2]))], 0)
    assert driver.assemble(full, 4, 1e-5)["cx"][:, 0].tolist() == [1.0, 2.0]
