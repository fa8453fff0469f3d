"""NVSwitch multicast (NVLS) slabs: one physical allocation per GPU bound to a shared multicast object.

A slab created here has two device addresses on every rank:

* ``local_ptr`` — ordinary (unicast) mapping of this GPU's own physical memory;
* ``mc_ptr``    — the multicast mapping.  ``multimem.ld_reduce`` on it returns the element-wise sum of *all* ranks'
  copies, computed inside the NVSwitch; ``multimem.st`` on it writes all ranks' copies at once.

``fedavg_finish_kernel`` uses the first for the cross-GPU FedAvg reduction of the partial sums and the second to
publish the new global model, so the server step moves ``P/world`` floats per GPU over NVLink instead of ``P`` and the
next round's broadcast reads local HBM only.

Plumbing (driver VMM API through cuda-python; no NCCL, no torch symmetric memory):
rank 0 ``cuMulticastCreate`` -> POSIX fd -> every peer over an abstract unix socket (SCM_RIGHTS) ->
``cuMulticastAddDevice`` on all ranks -> barrier -> ``cuMemCreate`` + ``cuMulticastBindMem`` -> map both handles.

Everything is optional: :func:`create` returns ``None`` (after all ranks agreed) when the GPU, driver, fabric manager
or container does not support multicast, and the callers fall back to the CUDA-IPC peer-pointer path
(parallel/comm.py).  Reference parity: replaces the S3/MinIO model upload + download of one round
(ols_core/taskMgr/utils/utils_run_task.py:327-397).
"""
from __future__ import annotations

import os
import socket
import time
from dataclasses import dataclass
from typing import List, Optional

import torch
import torch.distributed as dist

ENV = "OLSB_NVLS"            # "0" never, "1"/"auto" try when world > 1
_last_error: Optional[str] = None


def enabled(world: int) -> bool:
    return world > 1 and os.environ.get(ENV, "auto").lower() not in ("0", "off", "false", "no")


def min_bytes() -> int:
    """Slabs smaller than this stay on the CUDA-IPC path (``OLSB_NVLS_MIN_BYTES``, default 4 MiB; ``OLSB_NVLS=1`` forces
    multicast for every size)."""
    if os.environ.get(ENV, "auto") == "1":
        return 0
    return int(os.environ.get("OLSB_NVLS_MIN_BYTES", 4 << 20))


def last_error() -> Optional[str]:
    return _last_error


@dataclass
class MulticastSlab:
    nbytes: int              # usable bytes (request rounded up to the multicast granularity)
    local_ptr: int
    mc_ptr: int
    _handles: tuple          # (driver module, mc handle, mem handle, device)

    def close(self) -> None:
        cu, mc, mem, dev = self._handles
        try:
            cu.cuMemUnmap(self.mc_ptr, self.nbytes)
            cu.cuMemAddressFree(self.mc_ptr, self.nbytes)
            cu.cuMemUnmap(self.local_ptr, self.nbytes)
            cu.cuMemAddressFree(self.local_ptr, self.nbytes)
            cu.cuMulticastUnbind(mc, dev, 0, self.nbytes)
            cu.cuMemRelease(mem)
            cu.cuMemRelease(mc)
        except Exception:  # noqa: BLE001 — teardown is best effort (process exit frees everything)
            pass


class _Fail(RuntimeError):
    pass


def _ck(ret, what: str):
    """cuda-python returns (CUresult, values...)."""
    err = ret[0]
    if int(err) != 0:
        raise _Fail(f"{what}: CUresult {int(err)} ({getattr(err, 'name', err)})")
    return ret[1] if len(ret) == 2 else ret[1:]


def _sock_name(tag: str) -> bytes:
    return ("\0olsb-nvls-%s-%s" % (os.environ.get("MASTER_PORT", "0"), tag)).encode()


def _serve_fd(fd: int, npeers: int, tag: str, timeout: float) -> None:
    srv = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
    srv.bind(_sock_name(tag))
    srv.listen(npeers)
    srv.settimeout(timeout)
    try:
        for _ in range(npeers):
            conn, _a = srv.accept()
            with conn:
                socket.send_fds(conn, [b"fd"], [fd])
    finally:
        srv.close()


def _fetch_fd(tag: str, timeout: float) -> int:
    t_end = time.monotonic() + timeout
    while True:
        s = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
        try:
            s.connect(_sock_name(tag))
            s.settimeout(timeout)
            _msg, fds, _flags, _addr = socket.recv_fds(s, 16, 1)
            if not fds:
                raise _Fail("no descriptor received")
            return fds[0]
        except (FileNotFoundError, ConnectionRefusedError):
            if time.monotonic() > t_end:
                raise _Fail("multicast handle server not reachable") from None
            time.sleep(0.02)
        finally:
            s.close()


def _all_ok(ok: bool, group) -> bool:
    flags: List[Optional[bool]] = [None] * dist.get_world_size(group)
    dist.all_gather_object(flags, bool(ok), group=group)
    return all(flags)


_seq = 0


def create(nbytes: int, device: torch.device, group=None, timeout: float = 20.0) -> Optional[MulticastSlab]:
    """Collective over ``group``.  Returns a zero-filled slab or ``None`` on every rank (never a mix)."""
    global _last_error, _seq
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    _seq += 1
    tag = f"{_seq}"
    slab: Optional[MulticastSlab] = None
    stage = "import"
    cu = mc = mem = None
    dev_ord = device.index if device.index is not None else torch.cuda.current_device()
    ok = True
    try:
        from cuda.bindings import driver as cu          # type: ignore
        torch.cuda.init()
        torch.empty(1, device=device)                      # make sure the primary context exists and is current
        _ck(cu.cuInit(0), "cuInit")
        dev = _ck(cu.cuDeviceGet(dev_ord), "cuDeviceGet")
        sup = _ck(cu.cuDeviceGetAttribute(cu.CUdevice_attribute.CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev), "attr")
        if not int(sup):
            raise _Fail("CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED = 0")
        HT = cu.CUmemAllocationHandleType.CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR
        prop = cu.CUmulticastObjectProp()
        prop.numDevices = world
        prop.handleTypes = HT
        prop.flags = 0
        prop.size = max(int(nbytes), 1)
        stage = "granularity"
        gran = int(_ck(cu.cuMulticastGetGranularity(
            prop, cu.CUmulticastGranularity_flags.CU_MULTICAST_GRANULARITY_RECOMMENDED), "cuMulticastGetGranularity"))
        aprop = cu.CUmemAllocationProp()
        aprop.type = cu.CUmemAllocationType.CU_MEM_ALLOCATION_TYPE_PINNED
        aprop.location.type = cu.CUmemLocationType.CU_MEM_LOCATION_TYPE_DEVICE
        aprop.location.id = dev_ord
        aprop.requestedHandleTypes = HT
        mgran = int(_ck(cu.cuMemGetAllocationGranularity(
            aprop, cu.CUmemAllocationGranularity_flags.CU_MEM_ALLOC_GRANULARITY_RECOMMENDED), "cuMemGetAllocationGranularity"))
        gran = max(gran, mgran)
        size = (int(nbytes) + gran - 1) // gran * gran
        prop.size = size
    except Exception as e:  # noqa: BLE001
        ok, _last_error = False, f"{stage}: {e}"
    if not _all_ok(ok, group):
        return None
    # ---- multicast object: created by rank 0, handed to the peers as a file descriptor
    try:
        stage = "create"
        if rank == 0:
            mc = _ck(cu.cuMulticastCreate(prop), "cuMulticastCreate")
            fd = int(_ck(cu.cuMemExportToShareableHandle(mc, HT, 0), "cuMemExportToShareableHandle"))
            _serve_fd(fd, world - 1, tag, timeout)
            os.close(fd)
        else:
            fd = _fetch_fd(tag, timeout)
            mc = _ck(cu.cuMemImportFromShareableHandle(fd, HT), "cuMemImportFromShareableHandle")
            os.close(fd)
        stage = "add_device"
        _ck(cu.cuMulticastAddDevice(mc, dev), "cuMulticastAddDevice")
    except Exception as e:  # noqa: BLE001
        ok, _last_error = False, f"{stage}: {e}"
    if not _all_ok(ok, group):          # also the "every device was added" barrier required before binding
        return None
    try:
        stage = "mem_create"
        mem = _ck(cu.cuMemCreate(size, aprop, 0), "cuMemCreate")
        stage = "bind"
        _ck(cu.cuMulticastBindMem(mc, 0, mem, 0, size, 0), "cuMulticastBindMem")
        acc = cu.CUmemAccessDesc()
        acc.location.type = cu.CUmemLocationType.CU_MEM_LOCATION_TYPE_DEVICE
        acc.location.id = dev_ord
        acc.flags = cu.CUmemAccess_flags.CU_MEM_ACCESS_FLAGS_PROT_READWRITE
        stage = "map_local"
        va = _ck(cu.cuMemAddressReserve(size, gran, 0, 0), "cuMemAddressReserve")
        _ck(cu.cuMemMap(va, size, 0, mem, 0), "cuMemMap")
        _ck(cu.cuMemSetAccess(va, size, [acc], 1), "cuMemSetAccess")
        stage = "map_multicast"
        mva = _ck(cu.cuMemAddressReserve(size, gran, 0, 0), "cuMemAddressReserve(mc)")
        _ck(cu.cuMemMap(mva, size, 0, mc, 0), "cuMemMap(mc)")
        _ck(cu.cuMemSetAccess(mva, size, [acc], 1), "cuMemSetAccess(mc)")
        slab = MulticastSlab(size, int(va), int(mva), (cu, mc, mem, dev))
        from .comm import wrap_device_ptr
        wrap_device_ptr(slab.local_ptr, size, torch.uint8, (size,), device).zero_()
        torch.cuda.synchronize(device)
    except Exception as e:  # noqa: BLE001
        ok, _last_error = False, f"{stage}: {e}"
    if not _all_ok(ok, group):          # doubles as "every rank bound its memory": the slab is usable after this
        if slab is not None:
            slab.close()
        return None
    return slab
