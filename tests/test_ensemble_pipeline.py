"""The ensemble driver's three-stage pipeline (run.run_replicas_on_device) on stand-in members: which
thread runs which stage, in what order, what overlaps, and what happens when a stage fails.  No GPU."""
import threading
import time

import pytest

from multimm_b200 import run


class FakeMember:
    log = []          # (event, replica, thread name, time)
    lock = threading.Lock()
    fail = {}         # replica -> stage that raises

    def __init__(self, i):
        self.i = i
        self.timings = {"ingest_s": 0.0}
        self.report = {"iterations": 1, "converged": 1}
        self.closed = False

    def _stage(self, name, seconds):
        with FakeMember.lock:
            FakeMember.log.append((name + ":start", self.i, threading.current_thread().name, time.time()))
        if FakeMember.fail.get(self.i) == name:
            raise RuntimeError(f"{name} of member {self.i} failed")
        time.sleep(seconds)
        with FakeMember.lock:
            FakeMember.log.append((name + ":end", self.i, threading.current_thread().name, time.time()))

    def prepare(self):
        self._stage("prepare", 0.05)

    def compute(self):
        self._stage("compute", 0.15)

    def finish(self):
        self._stage("finish", 0.05)
        return self.report

    def close(self):
        self.closed = True


@pytest.fixture
def fake(monkeypatch):
    FakeMember.log, FakeMember.fail = [], {}
    made = {}

    def build(params, i, run_path, device):
        with FakeMember.lock:
            FakeMember.log.append(("build", i, threading.current_thread().name, time.time()))
        made[i] = FakeMember(i)
        return made[i]

    monkeypatch.setattr(run, "build_replica", build)
    monkeypatch.delenv("MMM_ENSEMBLE_PIPELINE", raising=False)
    return made


def _times(event, i):
    return [t for e, r, _, t in FakeMember.log if e == event and r == i][0]


def test_stages_run_on_their_threads_and_overlap(fake):
    got = []
    t0 = time.time()
    run.run_replicas_on_device({}, {i: f"/nowhere/{i}" for i in range(5)}, range(5), 0, False, got.append)
    wall = time.time() - t0
    assert [kind for kind, _ in got] == ["ok"] * 5 and [p["replica"] for _, p in got] == list(range(5))
    main = threading.current_thread().name
    by_stage = {}
    for e, r, th, _ in FakeMember.log:
        by_stage.setdefault(e.split(":")[0], set()).add(th)
    assert by_stage["compute"] == {main}
    assert len(by_stage["prepare"]) == 1 and by_stage["prepare"] == by_stage["build"] and main not in by_stage["prepare"]
    assert len(by_stage["finish"]) == 1 and by_stage["finish"].isdisjoint(by_stage["prepare"] | {main})
    # numpy's global random stream: build (seeds it) and prepare (draws on) of a member are adjacent, in member order
    seq = [(e, r) for e, r, _, _ in FakeMember.log if e in ("build", "prepare:start", "prepare:end")]
    assert seq == [x for i in range(5) for x in (("build", i), ("prepare:start", i), ("prepare:end", i))]
    # member k + 1 is prepared and member k - 1 written out while member k computes
    for k in range(1, 4):
        assert _times("prepare:end", k + 1) <= _times("compute:end", k)
        assert _times("finish:start", k - 1) <= _times("compute:end", k)
    # the GPU stage runs back to back: 5 x 0.15 s of compute, one prepare before, one finish after
    assert wall < 5 * 0.15 + 0.05 + 0.05 + 0.25, wall
    assert all(m.closed for m in fake.values())
    assert all({"compute_s", "finish_s", "seconds", "prepare_s"} <= set(p) for _, p in got)


@pytest.mark.parametrize("stage", ["prepare", "compute", "finish"])
def test_a_failing_stage_is_reported_and_the_others_go_on(fake, stage):
    FakeMember.fail = {2: stage}
    got = []
    run.run_replicas_on_device({}, {i: f"/nowhere/{i}" for i in range(4)}, range(4), 0, False, got.append)
    ok = sorted(p["replica"] for kind, p in got if kind == "ok")
    bad = [p for kind, p in got if kind == "error"]
    assert ok == [0, 1, 3] and len(bad) == 1 and bad[0]["replica"] == 2 and stage in bad[0]["error"]
    assert all(m.closed for m in fake.values())  # the failed member's engine is released too


def test_pipeline_can_be_switched_off(fake, monkeypatch):
    monkeypatch.setenv("MMM_ENSEMBLE_PIPELINE", "0")
    got = []
    run.run_replicas_on_device({}, {i: f"/nowhere/{i}" for i in range(3)}, range(3), 0, False, got.append)
    assert [p["replica"] for kind, p in got if kind == "ok"] == [0, 1, 2]
    assert {th for _, _, th, _ in FakeMember.log} == {threading.current_thread().name}
    order = [(e, r) for e, r, _, _ in FakeMember.log if e.endswith(":start")]
    assert order == [(s + ":start", i) for i in range(3) for s in ("prepare", "compute", "finish")]
