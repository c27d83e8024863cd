"""pylibwholegraph/torch/initialize.py of the reference (loaded unchanged through compat/) next to this repo's over fakes of
the binding, torch.cuda and torch.distributed: the same library / torch calls with the same arguments, the same environment
variables, the same world information handed to the communicator helpers, the same return values.

Documented differences: MASTER_ADDR defaults to 127.0.0.1 here (the reference says "localhost", which need not resolve in
a container); an already initialised process group is reused here; `backend="gloo"` is accepted for CPU-only control-plane
tests.  CPU only."""
import os
import types

import pytest
import torch

import wholegraph_b200.binding as wmb
from wholegraph_b200.torch import comm as comm_mod
from wholegraph_b200.torch import initialize as our_mod

REF = "/root/reference/python/pylibwholegraph/pylibwholegraph/torch/initialize.py"
pytestmark = pytest.mark.skipif(not os.path.exists(REF), reason="reference tree not present")


@pytest.fixture()
def layers(monkeypatch):
    from compat_loader import load_reference_file
    ref_mod = load_reference_file(REF, "_reference_initialize", package="pylibwholegraph.torch")
    log = []
    state = {"pg": False}
    monkeypatch.setattr(wmb, "init", lambda flags, level: log.append(("wmb.init", flags, int(level))))
    monkeypatch.setattr(wmb, "finalize", lambda: log.append(("wmb.finalize",)))
    monkeypatch.setattr(torch, "set_num_threads", lambda n: log.append(("set_num_threads", n)))
    monkeypatch.setattr(torch.cuda, "set_device", lambda d: log.append(("cuda.set_device", d)))

    def init_pg(backend=None, init_method=None, **kw):
        log.append(("init_process_group", backend, init_method, os.environ.get("RANK"), os.environ.get("WORLD_SIZE"),
                    os.environ.get("MASTER_PORT")))
        state["pg"] = True

    def destroy_pg():
        log.append(("destroy_process_group",))
        state["pg"] = False

    monkeypatch.setattr(torch.distributed, "init_process_group", init_pg)
    monkeypatch.setattr(torch.distributed, "destroy_process_group", destroy_pg)
    monkeypatch.setattr(torch.distributed, "is_initialized", lambda: state["pg"])
    for mod in (our_mod, ref_mod):
        monkeypatch.setattr(mod, "set_world_info", lambda *a: log.append(("set_world_info",) + a))
        monkeypatch.setattr(mod, "get_global_communicator", lambda backend="nccl": ("global", backend))
        monkeypatch.setattr(mod, "get_local_node_communicator", lambda: "node")
        monkeypatch.setattr(mod, "reset_communicators", lambda: log.append(("reset_communicators",)))
    for var in ("RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT"):
        monkeypatch.delenv(var, raising=False)
    return types.SimpleNamespace(ours=our_mod, ref=ref_mod, log=log, state=state)


def _run(layers, mod, script, env):
    del layers.log[:]
    layers.state["pg"] = False
    for var in ("RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT"):
        os.environ.pop(var, None)
    os.environ.update(env)
    result = script(mod)
    return result, sorted(layers.log, key=repr), {v: os.environ.get(v) for v in ("RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT")}


@pytest.mark.parametrize("env", [{}, {"MASTER_ADDR": "10.0.0.5", "MASTER_PORT": "4000"}, {"MASTER_PORT": "777"}])
def test_same_calls_same_environment(layers, env, capsys):
    scripts = {
        "init": lambda m: m.init(3, 8, 3, 8, "debug"),
        "init_torch_env": lambda m: m.init_torch_env(5, 16, 5, 8),
        "init + comm": lambda m: m.init_torch_env_and_create_wm_comm(1, 2, 1, 2, "nccl", "warn"),
        "finalize with a process group": lambda m: (m.init_torch_env(0, 1, 0, 1), m.finalize())[1],
        "finalize without one": lambda m: m.finalize(),
        "bad log level": lambda m: _raises(lambda: m.init(0, 1, 0, 1, "chatty")),
    }
    for name, script in scripts.items():
        ours = _run(layers, layers.ours, script, env)
        ref = _run(layers, layers.ref, script, env)
        if "MASTER_ADDR" not in env and ref[2]["MASTER_ADDR"] == "localhost":
            ref[2]["MASTER_ADDR"] = "127.0.0.1"       # the documented difference
        assert ours == ref, (name, ours, ref)         # same calls (as a multiset: the order of independent calls may differ)
    capsys.readouterr()


def _raises(fn):
    try:
        fn()
        return None
    except Exception as e:
        return type(e).__name__


def test_documented_extras(layers):
    """gloo is accepted and does not touch the GPU; an existing process group is reused."""
    _run(layers, layers.ours, lambda m: m.init_torch_env(0, 2, 0, 2, backend="gloo"), {})
    assert not any(c[0] == "cuda.set_device" for c in layers.log) and ("init_process_group", "gloo", "env://", "0", "2", "12335") in layers.log
    del layers.log[:]
    layers.ours.init_torch_env(0, 2, 0, 2)   # the group created above is still there
    assert not any(c[0] == "init_process_group" for c in layers.log)
