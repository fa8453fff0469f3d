"""Multi-GPU: the fused P2P broadcast / reduce kernels over N ranks must reproduce the single-rank model."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("opt", ["avg", "adam"])
def test_two_ranks_match_single_rank(opt):
    env = dict(os.environ, OPT=opt, CLIENTS="40")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533" if opt == "avg" else "29534", os.path.join(ROOT, "tools", "mgpu_check.py")]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600, cwd=ROOT)
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert lines, out.stderr[-2000:]
    res = json.loads(lines[-1])
    assert res["world"] == 2 and res["ok"], res


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("model,opt", [("cnn2", "avg"), ("cnn2", "adam"), ("mlp2", "momentum"), ("resnet18", "avg"), ("bert_mini", "avg")])
def test_native_ranks_match_the_nccl_oracle(model, opt):
    """native P2P broadcast / reduce + hand-written local step on 2 ranks vs the fp32 torch backend over NCCL all-reduce"""
    env = dict(os.environ, MODEL=model, OPT=opt)
    port = 29540 + abs(hash((model, opt))) % 40
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tools", "mgpu_oracle.py")]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900, cwd=ROOT)
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert lines, out.stderr[-2000:]
    res = json.loads(lines[-1])
    assert res["world"] == 2 and res["ok"], res


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_ipc_peer_pointer_fallback_matches_the_nccl_oracle():
    """the same check with the NVSwitch multicast slabs disabled: the CUDA-IPC peer-pointer reduce / broadcast path"""
    env = dict(os.environ, MODEL="cnn2", OPT="avg", OLSB_NVLS="0")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29597", os.path.join(ROOT, "tools", "mgpu_oracle.py")]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900, cwd=ROOT)
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert lines, out.stderr[-2000:]
    res = json.loads(lines[-1])
    assert res["world"] == 2 and res["ok"] and not res["nvls_multicast"], res


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_lost_peer_raises_on_the_survivor_and_keeps_its_context():
    """a rank that stops taking part mid-task: the survivor's reduce barrier gives up, PeerLostError surfaces on the host,
    the global model is untouched and the CUDA context still works (no device trap)"""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29598", os.path.join(ROOT, "tools", "peer_loss_check.py")]
    out = subprocess.run(cmd, env=dict(os.environ, LOST_SLEEP="8"), capture_output=True, text=True, timeout=600, cwd=ROOT)
    recs = [json.loads(l) for l in out.stdout.splitlines() if l.startswith("{")]
    surv = [r for r in recs if r.get("role") == "survivor"]
    assert surv and all(r["ok"] for r in surv), (recs, out.stderr[-1500:])
