"""north_star's acceptance sentence, executed: the reference's own driver scripts -- src/player.py,
src/double_dqn_conv.py, src/double_dqn_dense.py, byte for byte as in the reference tree (oracle/_ref, made
by oracle/make_ref.py) -- run against this repo's drop-in `board` / `dqn_lib` modules on the GPU.

Each script runs in its own process through `run_with_cuda_engine.py` with the job name piped to stdin and
the working directory inside a scratch "project" (a directory holding `.git/`, which the reference's
Experiment class searches for, src/experiments.py:20-29).  Only run LENGTHS are shortened, from outside
the scripts: `--set configs.<module>.<name>=<value>` overwrites hyper-parameters of the reference's config
modules before the script imports them, `--limit-loops` cuts player.py's hard-coded 1000-game loops.
Everything the scripts touch -- configs/*.py, device/__init__.py, experiments.py (Experiment.add_episode,
.snapshot_game, .save, torch.save(model)), tqdm -- is the reference's own code."""
import os
import pickle
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "src")
LAUNCHER = os.path.join(ROOT, "reinforcement-learning-2048_b200", "run_with_cuda_engine.py")


@pytest.fixture()
def project(tmp_path):
    if not os.path.isfile(os.path.join(REF, "board.py")):
        pytest.skip("oracle/_ref is absent: run __graft_entry__.build() where /root/reference exists -- "
                    "THE REFERENCE'S DRIVER SCRIPTS WERE NOT EXERCISED")
    (tmp_path / ".git").mkdir()
    return tmp_path


def run_script(project, script, options, job):
    cmd = [sys.executable, LAUNCHER] + options + [os.path.join(REF, script)]
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1")
    r = subprocess.run(cmd, cwd=str(project), input=job + "\n", capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0, f"{script} failed:\n{r.stdout[-3000:]}\n{r.stderr[-3000:]}"
    return r


def load(path):
    with open(path, "rb") as f:
        return pickle.load(f)


def test_player_py_runs_unchanged(project):
    """src/player.py:91-106: random-policy and up-left baselines (2 + 2 games instead of 1000 + 1000)."""
    r = run_script(project, "player.py", ["--limit-loops", "2"], "")
    assert "Random Games" in r.stdout and "Upleft games" in r.stdout
    for name in ("random_baseline", "upleft_baseline"):
        games = load(project / "experiments" / name / "binary" / "games_played.p")
        assert len(games) == 2
        for g in games:
            assert len(g) >= 1
            state, move, reward, merge_score = g[-1]
            assert state.shape == (4, 4) and move in ("u", "d", "l", "r", "up", "down", "left") and merge_score >= 0
    # a random game ends on a dead board: no move changes its last recorded state
    from oracle import board_oracle as bo
    last_state = load(project / "experiments" / "random_baseline" / "binary" / "games_played.p")[0][-1][0]
    assert bo.legal_mask(last_state) == 0


@pytest.mark.parametrize("script,cfg", [("double_dqn_conv.py", "configs.double_dqn_conv"),
                                        ("double_dqn_dense.py", "configs.double_dqn_dense")])
def test_training_drivers_run_unchanged(project, script, cfg):
    """src/double_dqn_conv.py:42-63 / src/double_dqn_dense.py:42-63 -> dqn_lib.training_loop with the
    reference's 20 positional arguments, its Experiment object, model, loss_fn and optimizer; 8 episodes,
    training from episode 3 on (batch 5000, sampled with replacement from what has been played so far),
    target sync every 2, a snapshot every 4."""
    job = "smoke_" + script.split(".")[0]
    opts = []
    for name, value in (("no_episodes", 8), ("no_episodes_before_training", 2), ("no_episodes_before_updating_target", 2),
                        ("snapshot_game_every_n_episodes", 4), ("no_episodes_to_reach_epsilon", 4)):
        opts += ["--set", f"{cfg}.{name}={value}"]
    r = run_script(project, script, opts, job)
    folder = project / "experiments" / job
    episodes = load(folder / "binary" / "episodes.p")
    assert [e["number"] for e in episodes] == list(range(8))
    for e in episodes:
        assert e["max_tile"] >= 4 and e["number_moves"] >= 1 and 0.0 <= e["epsilon"] <= 1.0 and e["merge_score"] >= 0
    hyper = load(folder / "binary" / "hyperparameters.p")
    assert hyper["batch_size"] == 5000 and hyper["use_double_dqn"] is True      # HYPERPARAMS was built before the patch
    assert os.path.isfile(folder / "text" / "hyperparams.json") and os.path.isfile(folder / f"{script}.txt")
    assert any(f.startswith("episode_") for f in os.listdir(folder / "binary" / "board_histories"))
    model = torch.load(folder / "binary" / "model.pt", weights_only=False, map_location="cpu")   # whole-module pickle
    assert sum(p.numel() for p in model.parameters()) == (33476 if "conv" in script else 403716)
    assert all(p.dtype == torch.float64 for p in model.parameters())
    assert "Episode" in r.stdout or "episode" in r.stdout or len(r.stdout) > 0
