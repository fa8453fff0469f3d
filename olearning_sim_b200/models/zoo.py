"""Model families of the FL simulator as *functional* specs over a packed flat parameter vector.

Every family is described by an ordered list of named parameter shapes and a pure ``apply(params_dict, x)``
function, so a whole population of clients is one ``[C, P]`` tensor and ``torch.func.vmap(grad(loss))`` gives the
per-client gradients in one call (the generic/library backend; the CNN-2 family additionally has the hand-written
sm_100a path in ``engine/cnn2_native.py``).  ``torch.nn`` twins are available through ``to_module`` for the
baseline and for checkpoint interchange.

Families (BASELINE.json configs): ``mlp2`` (#1), ``cnn2`` (#2, #4), ``resnet18`` (#3, GroupNorm instead of
BatchNorm — the FL-standard choice, batch statistics do not average across non-IID clients), ``bert_mini`` (#5).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Callable, Dict, List, Tuple

import torch
import torch.nn.functional as F


@dataclass
class ModelSpec:
    name: str
    shapes: List[Tuple[str, Tuple[int, ...]]]
    apply: Callable[[Dict[str, torch.Tensor], torch.Tensor], torch.Tensor]
    input_shape: Tuple[int, ...]
    num_classes: int
    input_dtype: torch.dtype = torch.float32

    @property
    def num_params(self) -> int:
        return sum(math.prod(s) for _, s in self.shapes)

    @property
    def padded_params(self) -> int:
        return (self.num_params + 3) // 4 * 4

    def unflatten(self, flat: torch.Tensor) -> Dict[str, torch.Tensor]:
        out, o = {}, 0
        for n, s in self.shapes:
            k = math.prod(s)
            out[n] = flat[o:o + k].reshape(s)
            o += k
        return out

    def flatten(self, params: Dict[str, torch.Tensor]) -> torch.Tensor:
        flat = torch.cat([params[n].reshape(-1).float() for n, _ in self.shapes])
        pad = self.padded_params - flat.numel()
        return F.pad(flat, (0, pad)) if pad else flat

    def init(self, seed: int = 0) -> torch.Tensor:
        g = torch.Generator().manual_seed(seed)
        parts = []
        for n, s in self.shapes:
            if n.endswith(".bias") or n.endswith(".beta"):
                parts.append(torch.zeros(s))
            elif n.endswith(".gamma"):
                parts.append(torch.ones(s))
            elif n.endswith(".emb"):
                parts.append(torch.randn(s, generator=g) * 0.02)
            else:
                fan_in = math.prod(s[1:]) if len(s) > 1 else s[0]
                bound = 1.0 / math.sqrt(fan_in)          # torch.nn default (kaiming_uniform with a=sqrt(5))
                parts.append((torch.rand(s, generator=g) * 2 - 1) * bound)
        return self.flatten({n: p for (n, _), p in zip(self.shapes, parts)})

    def loss(self, flat: torch.Tensor, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        logits = self.apply(self.unflatten(flat), x)
        return F.cross_entropy(logits, y.long())


# ---- MLP-2 ------------------------------------------------------------------------------------------
def mlp2(in_dim: int = 784, hidden: int = 200, num_classes: int = 10) -> ModelSpec:
    def apply(p, x):
        h = F.relu(F.linear(x.reshape(x.shape[0], -1), p["fc1.weight"], p["fc1.bias"]))
        return F.linear(h, p["fc2.weight"], p["fc2.bias"])
    return ModelSpec("mlp2", [("fc1.weight", (hidden, in_dim)), ("fc1.bias", (hidden,)),
                              ("fc2.weight", (num_classes, hidden)), ("fc2.bias", (num_classes,))],
                     apply, (in_dim,), num_classes)


def mlp(in_dim: int = 784, hidden=(256, 128), num_classes: int = 10) -> ModelSpec:
    """MLP with any number of hidden layers (fc1 .. fcN, ReLU between) — native on the GPU through the per-client GEMM
    family, like ``mlp2``."""
    dims = [in_dim] + [int(h) for h in hidden] + [num_classes]
    n = len(dims) - 1

    def apply(p, x):
        h = x.reshape(x.shape[0], -1)
        for l in range(1, n + 1):
            h = F.linear(h, p[f"fc{l}.weight"], p[f"fc{l}.bias"])
            if l < n:
                h = F.relu(h)
        return h
    shapes = []
    for l in range(1, n + 1):
        shapes += [(f"fc{l}.weight", (dims[l], dims[l - 1])), (f"fc{l}.bias", (dims[l],))]
    return ModelSpec("mlp", shapes, apply, (in_dim,), num_classes)


# ---- CNN-2 (CIFAR-shaped) -----------------------------------------------------------------------------
def cnn2() -> ModelSpec:
    def apply(p, x):  # x [B,3,32,32]
        h = F.max_pool2d(F.relu(F.conv2d(x, p["conv1.weight"], p["conv1.bias"], padding=2)), 2)
        h = F.max_pool2d(F.relu(F.conv2d(h, p["conv2.weight"], p["conv2.bias"], padding=2)), 2)
        return F.linear(h.reshape(h.shape[0], -1), p["fc.weight"], p["fc.bias"])
    return ModelSpec("cnn2", [("conv1.weight", (32, 3, 5, 5)), ("conv1.bias", (32,)),
                              ("conv2.weight", (64, 32, 5, 5)), ("conv2.bias", (64,)),
                              ("fc.weight", (10, 4096)), ("fc.bias", (10,))], apply, (3, 32, 32), 10)


# ---- ResNet-18 (GroupNorm, CIFAR stem) ----------------------------------------------------------------
def resnet18(num_classes: int = 10, groups: int = 2, width: int = 64) -> ModelSpec:
    shapes: List[Tuple[str, Tuple[int, ...]]] = [("stem.weight", (width, 3, 3, 3)), ("stem.gn.gamma", (width,)),
                                                ("stem.gn.beta", (width,))]
    cfg = []
    cin = width
    for stage, (cout, stride) in enumerate([(width, 1), (2 * width, 2), (4 * width, 2), (8 * width, 2)]):
        for blk in range(2):
            s = stride if blk == 0 else 1
            name = f"l{stage}.{blk}"
            shapes += [(f"{name}.c1.weight", (cout, cin, 3, 3)), (f"{name}.g1.gamma", (cout,)), (f"{name}.g1.beta", (cout,)),
                       (f"{name}.c2.weight", (cout, cout, 3, 3)), (f"{name}.g2.gamma", (cout,)), (f"{name}.g2.beta", (cout,))]
            down = s != 1 or cin != cout
            if down:
                shapes += [(f"{name}.ds.weight", (cout, cin, 1, 1)), (f"{name}.dg.gamma", (cout,)), (f"{name}.dg.beta", (cout,))]
            cfg.append((name, s, down))
            cin = cout
    shapes += [("fc.weight", (num_classes, cin)), ("fc.bias", (num_classes,))]

    def gn(x, p, n):
        return F.group_norm(x, groups, p[f"{n}.gamma"], p[f"{n}.beta"])

    def apply(p, x):
        h = F.relu(gn(F.conv2d(x, p["stem.weight"], padding=1), p, "stem.gn"))
        for name, s, down in cfg:
            idt = h
            o = F.relu(gn(F.conv2d(h, p[f"{name}.c1.weight"], stride=s, padding=1), p, f"{name}.g1"))
            o = gn(F.conv2d(o, p[f"{name}.c2.weight"], padding=1), p, f"{name}.g2")
            if down:
                idt = gn(F.conv2d(h, p[f"{name}.ds.weight"], stride=s), p, f"{name}.dg")
            h = F.relu(o + idt)
        h = F.adaptive_avg_pool2d(h, 1).flatten(1)
        return F.linear(h, p["fc.weight"], p["fc.bias"])

    return ModelSpec("resnet18", shapes, apply, (3, 32, 32), num_classes)


# ---- BERT-mini (4 layers, H=256, 4 heads) -----------------------------------------------------------------
def bert_mini(vocab: int = 30522, seq: int = 128, hidden: int = 256, layers: int = 4, heads: int = 4,
              num_classes: int = 2) -> ModelSpec:
    shapes: List[Tuple[str, Tuple[int, ...]]] = [("tok.emb", (vocab, hidden)), ("pos.emb", (seq, hidden)),
                                                ("emb.ln.gamma", (hidden,)), ("emb.ln.beta", (hidden,))]
    for l in range(layers):
        n = f"h{l}"
        shapes += [(f"{n}.qkv.weight", (3 * hidden, hidden)), (f"{n}.qkv.bias", (3 * hidden,)),
                   (f"{n}.o.weight", (hidden, hidden)), (f"{n}.o.bias", (hidden,)),
                   (f"{n}.ln1.gamma", (hidden,)), (f"{n}.ln1.beta", (hidden,)),
                   (f"{n}.ff1.weight", (4 * hidden, hidden)), (f"{n}.ff1.bias", (4 * hidden,)),
                   (f"{n}.ff2.weight", (hidden, 4 * hidden)), (f"{n}.ff2.bias", (hidden,)),
                   (f"{n}.ln2.gamma", (hidden,)), (f"{n}.ln2.beta", (hidden,))]
    shapes += [("cls.weight", (num_classes, hidden)), ("cls.bias", (num_classes,))]
    hd = hidden // heads

    def apply(p, ids):  # ids [B, S] int64
        b, s = ids.shape
        h = F.embedding(ids, p["tok.emb"]) + p["pos.emb"][:s][None]
        h = F.layer_norm(h, (hidden,), p["emb.ln.gamma"], p["emb.ln.beta"])
        for l in range(layers):
            n = f"h{l}"
            qkv = F.linear(h, p[f"{n}.qkv.weight"], p[f"{n}.qkv.bias"]).reshape(b, s, 3, heads, hd).permute(2, 0, 3, 1, 4)
            att = torch.softmax(qkv[0] @ qkv[1].transpose(-1, -2) / math.sqrt(hd), dim=-1) @ qkv[2]
            att = att.transpose(1, 2).reshape(b, s, hidden)
            h = F.layer_norm(h + F.linear(att, p[f"{n}.o.weight"], p[f"{n}.o.bias"]), (hidden,), p[f"{n}.ln1.gamma"], p[f"{n}.ln1.beta"])
            ff = F.linear(F.gelu(F.linear(h, p[f"{n}.ff1.weight"], p[f"{n}.ff1.bias"])), p[f"{n}.ff2.weight"], p[f"{n}.ff2.bias"])
            h = F.layer_norm(h + ff, (hidden,), p[f"{n}.ln2.gamma"], p[f"{n}.ln2.beta"])
        return F.linear(h[:, 0], p["cls.weight"], p["cls.bias"])

    return ModelSpec("bert_mini", shapes, apply, (seq,), num_classes, input_dtype=torch.int64)


REGISTRY: Dict[str, Callable[..., ModelSpec]] = {"mlp2": mlp2, "mlp": mlp, "cnn2": cnn2, "resnet18": resnet18, "bert_mini": bert_mini}


def get_model(name: str, **kw) -> ModelSpec:
    if name not in REGISTRY:
        raise KeyError(f"unknown model family {name!r}; available: {sorted(REGISTRY)}")
    return REGISTRY[name](**kw)


class _Functional(torch.nn.Module):
    """torch.nn twin of a spec (baseline / checkpoint interchange)."""

    def __init__(self, spec: ModelSpec, flat: torch.Tensor):
        super().__init__()
        self.spec = spec
        self.params = torch.nn.ParameterDict({n.replace(".", "__"): torch.nn.Parameter(v.clone())
                                              for n, v in spec.unflatten(flat).items()})

    def forward(self, x):
        return self.spec.apply({n.replace("__", "."): v for n, v in self.params.items()}, x)


def to_module(spec: ModelSpec, flat: torch.Tensor) -> torch.nn.Module:
    return _Functional(spec, flat)
