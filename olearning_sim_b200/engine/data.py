"""Client data shards for the built-in (GPU) operators — the reference's data contract on arrays.

Reference behaviour (ols_core/taskMgr/utils/utils_run_task.py:174-325): a dataset is a zip (fetched by FILE / HTTP / S3 /
MINIO) or an already-unpacked folder; ``data_split_type = false`` gives every simulated client the same folder,
``true`` requires exactly ``simulation_num`` sub-folders which are sorted and indexed by the client's ``split_index`` —
client *i* owns shard *i*, which is what makes the population non-IID.  The operator scripts then read whatever files
they like from that folder.

The built-in operators keep the folder contract and fix the file format: a shard folder (or the shared folder) holds
``data.npz`` with ``x`` (uint8 ``[S,32,32,3]`` images, or float ``[S, ...]`` features, or int64 token ids) and ``y`` (int
labels), or the pair ``x.npy`` / ``y.npy``.  ``load_client_shards`` returns exactly the arrays ``FedEngine.load_client_data``
takes: ``xs [C_local, shard, ...]``, ``ys [C_local, shard]`` for this rank's clients.
"""
from __future__ import annotations

import os
import zipfile
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np


def fetch_dataset(data: Dict[str, Any], work_dir: str, file_repo: Optional[Any] = None) -> str:
    """Resolve ``data_path`` to a local folder: download the zip when the transfer type is not FILE, unpack it once."""
    path = data.get("data_path", "")
    if not path:
        return ""
    ttype = str(data.get("data_transfer_type", "FILE")).upper()
    ttype = {"0": "FILE", "1": "HTTP", "2": "S3", "3": "MINIO"}.get(ttype, ttype)
    if not path.endswith(".zip"):
        return path                                        # an unpacked folder on a shared file system
    local_zip = path
    if ttype != "FILE":
        if file_repo is None:
            from ..utils.file_repo import get_file_repo
            file_repo = get_file_repo(ttype, data.get("repo_config"))
        os.makedirs(work_dir, exist_ok=True)
        local_zip = os.path.join(work_dir, os.path.basename(path))
        if not file_repo.download_if_missing(path, local_zip):
            raise FileNotFoundError(f"download of {path} ({ttype}) failed")
    if not os.path.exists(local_zip):
        raise FileNotFoundError(f"dataset zip {local_zip} does not exist")
    out = os.path.join(work_dir, os.path.splitext(os.path.basename(local_zip))[0])
    marker = os.path.join(out, ".unpacked")
    if not os.path.exists(marker):
        os.makedirs(out, exist_ok=True)
        with zipfile.ZipFile(local_zip) as z:
            for m in z.namelist():                          # refuse entries that escape the target folder
                dst = os.path.normpath(os.path.join(out, m))
                if not dst.startswith(os.path.normpath(out) + os.sep) and dst != os.path.normpath(out):
                    raise ValueError(f"zip entry escapes the dataset folder: {m}")
            z.extractall(out)
        open(marker, "w").close()
    # a zip that wraps everything in one top-level folder: descend into it
    entries = [e for e in os.listdir(out) if e != ".unpacked"]
    if len(entries) == 1 and os.path.isdir(os.path.join(out, entries[0])) and not _has_arrays(out):
        return os.path.join(out, entries[0])
    return out


def _has_arrays(folder: str) -> bool:
    return os.path.exists(os.path.join(folder, "data.npz")) or os.path.exists(os.path.join(folder, "x.npy"))


def _read_folder(folder: str) -> Tuple[np.ndarray, np.ndarray]:
    npz = os.path.join(folder, "data.npz")
    if os.path.exists(npz):
        with np.load(npz) as z:
            return np.asarray(z["x"]), np.asarray(z["y"])
    x, y = os.path.join(folder, "x.npy"), os.path.join(folder, "y.npy")
    if os.path.exists(x) and os.path.exists(y):
        return np.load(x), np.load(y)
    raise FileNotFoundError(f"{folder}: expected data.npz (x, y) or x.npy + y.npy")


def _fit(x: np.ndarray, y: np.ndarray, shard: int) -> Tuple[np.ndarray, np.ndarray]:
    """Exactly ``shard`` samples: truncate, or repeat cyclically when the folder holds fewer."""
    n = len(y)
    if n == 0:
        raise ValueError("empty shard")
    idx = np.arange(shard) % n
    return x[idx], y[idx]


def list_shard_folders(root: str) -> List[str]:
    """Sub-folders in the reference's order (plain ``sorted`` of the names, utils_run_task.py:248-263)."""
    return [os.path.join(root, d) for d in sorted(os.listdir(root)) if os.path.isdir(os.path.join(root, d))]


def load_client_shards(root: str, client_ids: Sequence[int], shard: int, split: bool, total_clients: int,
                       input_shape: Optional[Tuple[int, ...]] = None) -> Tuple[np.ndarray, np.ndarray]:
    """Arrays for this rank's clients.  ``split``: client i reads sub-folder i (there must be exactly ``total_clients``
    of them); otherwise every client reads the shared folder and takes the i-th contiguous block of it (wrapping)."""
    xs: List[np.ndarray] = []
    ys: List[np.ndarray] = []
    if split:
        folders = list_shard_folders(root)
        if len(folders) != total_clients:
            raise ValueError(f"data_split_type=true needs exactly {total_clients} sub-folders in {root}, found {len(folders)}")
        for cid in client_ids:
            x, y = _fit(*_read_folder(folders[int(cid)]), shard)
            xs.append(x); ys.append(y)
    else:
        x_all, y_all = _read_folder(root)
        n = len(y_all)
        for cid in client_ids:
            idx = (int(cid) * shard + np.arange(shard)) % n
            xs.append(x_all[idx]); ys.append(y_all[idx])
    X, Y = np.stack(xs), np.stack(ys).astype(np.int32)
    if input_shape is not None and tuple(X.shape[2:]) != tuple(input_shape):
        if int(np.prod(X.shape[2:])) != int(np.prod(input_shape)):
            raise ValueError(f"samples of shape {X.shape[2:]} do not match the model input {tuple(input_shape)}")
        X = X.reshape(X.shape[:2] + tuple(input_shape))
    return X, Y


# ---- building shard folders ------------------------------------------------------------------------------------------
def write_shards(out_dir: str, xs: Sequence[np.ndarray], ys: Sequence[np.ndarray], prefix: str = "client_") -> List[str]:
    """One sub-folder per client (``client_0000`` ...; zero-padded so the reference's plain ``sorted`` keeps the order),
    each holding ``data.npz`` — the layout ``load_client_shards(..., split=True)`` reads."""
    width = max(4, len(str(max(len(xs) - 1, 0))))
    out = []
    for i, (x, y) in enumerate(zip(xs, ys)):
        d = os.path.join(out_dir, f"{prefix}{i:0{width}d}")
        os.makedirs(d, exist_ok=True)
        np.savez(os.path.join(d, "data.npz"), x=np.asarray(x), y=np.asarray(y))
        out.append(d)
    return out


def dirichlet_partition(y: np.ndarray, num_clients: int, alpha: float, seed: int = 0, min_size: int = 1) -> List[np.ndarray]:
    """Label-skewed non-IID split (the usual FL benchmark protocol): for every class the samples are divided among the
    clients with proportions drawn from Dir(alpha).  Small alpha -> each client sees few classes; large alpha -> IID.
    Returns one index array per client (every sample is assigned exactly once)."""
    rng = np.random.default_rng(seed)
    y = np.asarray(y).reshape(-1)
    for _ in range(100):
        buckets: List[List[int]] = [[] for _ in range(num_clients)]
        for c in np.unique(y):
            idx = rng.permutation(np.flatnonzero(y == c))
            cuts = (np.cumsum(rng.dirichlet(np.full(num_clients, alpha))) * len(idx)).astype(int)[:-1]
            for b, part in zip(buckets, np.split(idx, cuts)):
                b.extend(part.tolist())
        if min(len(b) for b in buckets) >= min_size:
            break
    return [np.array(sorted(b), dtype=np.int64) for b in buckets]
