"""Vectorised, device-resident environment loop.

The reference's array API — `StepEnvNumpy.submit_instructions(...)` then `step()` then `level_2_data()`
(rust/src/step_sim_numpy.rs:233-275, :139-145, :351-368) — is the shape an RL loop uses, one env at a time with numpy
arrays on the host.  `VectorEnv` is that loop batched over `n_envs` envs with every array resident in DEVICE memory:
actions go in as a `[n_envs, rows_per_env]` block of packed instruction rows, the new order ids and the level-1 / level-2
observations come back as device arrays, and nothing crosses PCIe per step unless the caller asks for a host copy.

Buffers are exchanged through DLPack (`__dlpack__` / `__dlpack_device__`: `torch.from_dlpack(env.obs)`, `cupy.from_dlpack`,
`jax.dlpack`) and the CUDA array interface (`__cuda_array_interface__`, version 3), which torch, cupy and numba all
speak, so this module needs none of them: both give zero-copy views, and a torch / cupy array (anything with `__dlpack__`
or the CUDA array interface) can be passed to `step` directly.  numpy arrays are accepted too (copied host to device).
"""
from __future__ import annotations

import ctypes as C
import typing

import numpy as np

from . import abi
from .core import BatchedEnv

# ---- DLPack (dlpack.h, v0.8 ABI: DLManagedTensor in a PyCapsule named "dltensor"), pure ctypes ---------------------------
KDL_CUDA, KDL_UINT = 2, 1


class _DLDevice(C.Structure):
    _fields_ = [("device_type", C.c_int32), ("device_id", C.c_int32)]


class _DLDataType(C.Structure):
    _fields_ = [("code", C.c_uint8), ("bits", C.c_uint8), ("lanes", C.c_uint16)]


class _DLTensor(C.Structure):
    _fields_ = [("data", C.c_void_p), ("device", _DLDevice), ("ndim", C.c_int32), ("dtype", _DLDataType),
                ("shape", C.POINTER(C.c_int64)), ("strides", C.POINTER(C.c_int64)), ("byte_offset", C.c_uint64)]


class _DLManagedTensor(C.Structure):
    pass


_DLDeleter = C.CFUNCTYPE(None, C.POINTER(_DLManagedTensor))
_DLManagedTensor._fields_ = [("dl_tensor", _DLTensor), ("manager_ctx", C.c_void_p), ("deleter", _DLDeleter)]
_api = C.pythonapi
_DLTENSOR = b"dltensor"   # (PyCapsule keeps the POINTER to its name: the bytes object must outlive every capsule)
_api.PyCapsule_New.restype, _api.PyCapsule_New.argtypes = C.py_object, [C.c_void_p, C.c_char_p, C.c_void_p]
_api.PyCapsule_IsValid.restype, _api.PyCapsule_IsValid.argtypes = C.c_int, [C.py_object, C.c_char_p]
_api.PyCapsule_GetPointer.restype, _api.PyCapsule_GetPointer.argtypes = C.c_void_p, [C.py_object, C.c_char_p]
_api.PyCapsule_SetName.restype, _api.PyCapsule_SetName.argtypes = C.c_int, [C.py_object, C.c_char_p]
_LIVE_EXPORTS: dict = {}   # address of an exported DLManagedTensor -> everything that must outlive its consumer


@_DLDeleter
def _dl_deleter(mt_ptr):   # called by the consumer when it drops the tensor: release our bookkeeping (the env owns the memory)
    _LIVE_EXPORTS.pop(C.addressof(mt_ptr.contents), None)


def _dlpack_import(x) -> typing.Tuple[int, int, typing.Any]:
    """(device pointer, byte size, keep-alive object) of anything that exports DLPack; C-contiguous CUDA tensors only."""
    capsule = x.__dlpack__()
    if not _api.PyCapsule_IsValid(capsule, _DLTENSOR):
        raise ValueError("not a DLPack capsule")
    mt = C.cast(_api.PyCapsule_GetPointer(capsule, _DLTENSOR), C.POINTER(_DLManagedTensor)).contents
    t = mt.dl_tensor
    if t.device.device_type != KDL_CUDA:
        raise ValueError("DLPack tensor is not in CUDA device memory")
    shape = [t.shape[i] for i in range(t.ndim)]
    if t.strides:
        expect = 1
        for i in reversed(range(t.ndim)):
            if shape[i] != 1 and t.strides[i] != expect:
                raise ValueError("device arrays must be C-contiguous")
            expect *= shape[i]
    nbytes = int(np.prod(shape)) * (t.dtype.bits * t.dtype.lanes // 8)
    # the capsule is kept (un-renamed) together with its producer: its own destructor releases the tensor when we drop it
    return int(t.data) + int(t.byte_offset), nbytes, (capsule, x)

ACTION_DTYPE = abi.INSTR_DTYPE  # one row = one instruction: (t ignored, op_flags, order_id, price, vol, trader, aux)


class DeviceArray:
    """A C-contiguous array in device memory owned by a `BatchedEnv`; speaks `__cuda_array_interface__`."""

    def __init__(self, env: BatchedEnv, shape: typing.Tuple[int, ...], dtype):
        self._env, self.shape, self.dtype = env, tuple(shape), np.dtype(dtype)
        self.nbytes = int(np.prod(self.shape)) * self.dtype.itemsize
        self.ptr = env.device_alloc(self.nbytes)

    @property
    def __cuda_array_interface__(self):
        return {"shape": self.shape, "typestr": self.dtype.str, "data": (self.ptr, False), "version": 3, "strides": None}

    def __dlpack_device__(self):
        return (KDL_CUDA, self._env.device)

    def __dlpack__(self, stream=None, **_kw):
        """DLPack export (zero copy): `torch.from_dlpack(arr)`.  The memory stays owned by the env — keep the env alive while
        views exist.  `stream`: the consumer's stream; the producer side is the env's stream, which the caller orders
        against as for any other use of the env's outputs (`env.synchronize()` or stream semantics of `set_stream`)."""
        if self.dtype.kind != "u":
            raise TypeError("only unsigned integer arrays are exported")
        mt = _DLManagedTensor()
        shape = (C.c_int64 * len(self.shape))(*self.shape)
        mt.dl_tensor = _DLTensor(C.c_void_p(self.ptr), _DLDevice(KDL_CUDA, self._env.device), len(self.shape),
                                 _DLDataType(KDL_UINT, 8 * self.dtype.itemsize, 1), shape, None, 0)
        mt.manager_ctx, mt.deleter = None, _dl_deleter
        _LIVE_EXPORTS[C.addressof(mt)] = (mt, shape, self)
        # no capsule destructor: a capsule that is never consumed leaves its ~200-byte bookkeeping entry behind (a ctypes
        # callback must not be handed a capsule that is being deallocated); a consumed one is released by `_dl_deleter`
        return _api.PyCapsule_New(C.addressof(mt), _DLTENSOR, None)

    def numpy(self) -> np.ndarray:
        """Host copy (synchronises the env's stream)."""
        out = np.empty(self.shape, self.dtype)
        self._env.memcpy(out.ctypes.data, self.ptr, self.nbytes, 2)
        return out

    def copy_from_host(self, a: np.ndarray):
        a = np.ascontiguousarray(a)
        assert a.nbytes == self.nbytes, "size mismatch"
        self._env.memcpy(self.ptr, a.ctypes.data, self.nbytes, 1)

    def free(self):
        if self.ptr:
            self._env.device_free(self.ptr)
            self.ptr = 0


def _device_ptr(x, nbytes: int) -> typing.Optional[int]:
    cai = getattr(x, "__cuda_array_interface__", None)
    if cai is None:
        if hasattr(x, "__dlpack__") and not isinstance(x, np.ndarray):   # DLPack-only producers
            ptr, n, keep = _dlpack_import(x)
            if n != nbytes:
                raise ValueError(f"device array holds {n} bytes, expected {nbytes}")
            _device_ptr.keep = keep   # alive until the next call: the launch that reads it is already enqueued by then
            return ptr
        return None
    if cai.get("strides") is not None:
        raise ValueError("device arrays must be C-contiguous")
    n = int(np.prod(cai["shape"])) * np.dtype(cai["typestr"]).itemsize
    if n != nbytes:
        raise ValueError(f"device array holds {n} bytes, expected {nbytes}")
    return int(cai["data"][0])


def pack_actions(op, bid=None, vol=None, trader=None, price=None, order_id=None, market=None, has_price=None, has_vol=None) -> np.ndarray:
    """Columns -> packed instruction rows (host numpy, any shape).  `op`: abi.OP_NOOP / OP_NEW / OP_CANCEL / OP_MODIFY.
    NEW rows: `market=True` is a market order (`price=None` in the reference).  MODIFY rows: `has_price` / `has_vol` say
    which of `price` / `vol` is `Some` (default: both)."""
    op = np.asarray(op, dtype=np.uint32)
    a = np.zeros(op.shape, dtype=ACTION_DTYPE)

    def col(x, default=0):
        return np.broadcast_to(np.asarray(default if x is None else x), op.shape)

    f = op.copy()
    f |= np.where(col(bid).astype(bool) & (op == abi.OP_NEW), abi.F_BID, 0).astype(np.uint32)
    f |= np.where(col(market).astype(bool) & (op == abi.OP_NEW), abi.F_MARKET, 0).astype(np.uint32)
    mod = op == abi.OP_MODIFY
    f |= np.where(mod & col(has_price, 1).astype(bool), abi.F_HAS_PRICE, 0).astype(np.uint32)
    f |= np.where(mod & col(has_vol, 1).astype(bool), abi.F_HAS_VOL, 0).astype(np.uint32)
    a["op_flags"] = f
    oid = col(order_id).astype(np.uint64)
    a["order_id"] = np.where(oid > 0xFFFFFFFE, 0xFFFFFFFF, oid).astype(np.uint32)  # ids beyond u32 cannot exist: bad id
    a["price"], a["vol"], a["trader"] = col(price), col(vol), col(trader)
    return a


class VectorEnv:
    """`n_envs` lockstep `StepEnvNumpy(seed + env, start_time, tick_size, step_size, trading)` instances with a fixed action
    block of `rows_per_env` instruction rows per env and step (pad with OP_NOOP rows).

        env = VectorEnv(4096, rows_per_env=8, seed=0, start_time=0, tick_size=1, step_size=1000)
        obs = env.reset()                       # DeviceArray [n_envs, 45] u32, StepEnvNumpy.level_2_data layout
        obs, ids = env.step(actions)            # actions: [n_envs, rows_per_env] packed rows, device or host
        t = torch.as_tensor(obs, device="cuda") # zero-copy

    `ids[e, r]` is the order id row r created in env e, or `abi.NO_ID` (2^64 - 1) for cancel / modify / no-op rows
    (step_sim_numpy.rs:256-267).  Semantics per env are exactly `submit_instructions` + `step`: ids in row order, the
    step's queue shuffled with the env's own Xoroshiro stream, event i at `t + i`.  A NEW row with an off-tick limit
    price creates nothing (the reference raises ValueError there); the env is flagged and `check_errors` raises.

    Without background agents a step is ONE kernel launch on the env's stream (`env.set_stream(...)`) with no host
    synchronisation, so a loop of policy + `step` can be captured into a CUDA graph and replayed (tests/test_gpu_gym.py);
    steps that include the built-in agents refuse capture (their history staging is validated on the host).  The library's
    per-step history is a scratch ring for this class: `step` clears it (`bb_clear_history`) whenever `max_steps` records
    have accumulated; a caller replaying a captured step must do the same every `max_steps` replays."""

    def __init__(self, n_envs: int, rows_per_env: int, seed: int, start_time: int, tick_size: int, step_size: int,
                 trading: bool = True, *, level_1: bool = False, agents=None, agent_seed: int = 0, **kw):
        """`agents`: optional background population (a list of `core.random_group` / `momentum_group` / `noise_group`
        records).  With it every `step` is `{ agents.update(env); the action rows; env.step() }` in one launch
        (bb_run_agents_with_rows): the rows join the agents' instructions in the step's one shuffled queue (Philox key
        `agent_seed`).  Action rows may then be NEW, CANCEL or no-op (MODIFY is refused).  At one step per launch the
        general engine (no `price_window`) is the faster choice here: 92 us against 117 us per 4096-env step on the dense
        engine, whose per-launch agent-table rebuild only pays off inside the persistent multi-step kernel."""
        n_bg = sum(int(g["n_agents"]) for g in agents) if agents else 0
        kw.setdefault("max_queue", max(rows_per_env + 2 * n_bg, 16))
        if rows_per_env > kw["max_queue"]:
            raise ValueError("rows_per_env exceeds max_queue")
        self._agents, self._agent_seed = agents, agent_seed
        self._n_steps = 0
        self.n_envs, self.rows = n_envs, rows_per_env
        self.env = BatchedEnv(n_envs, seed, start_time, tick_size, step_size, trading, obs_words=abi.OBS_L1 if level_1 else abi.OBS_L2, **kw)
        self.obs_words = abi.OBS_L1 if level_1 else abi.OBS_L2
        self.obs = DeviceArray(self.env, (n_envs, self.obs_words), np.uint32)
        self.ids = DeviceArray(self.env, (n_envs, rows_per_env), np.uint64)
        self._actions = DeviceArray(self.env, (n_envs, rows_per_env), ACTION_DTYPE)  # staging for host-side actions
        self._offsets = DeviceArray(self.env, (n_envs + 1,), np.uint64)
        self._offsets.copy_from_host(np.arange(n_envs + 1, dtype=np.uint64) * np.uint64(rows_per_env))
        if agents:
            self.env.set_agents(agents)

    def close(self):
        for a in (self.obs, self.ids, self._actions, self._offsets):
            a.free()
        self.env.close()

    def _observe(self) -> DeviceArray:
        (self.env.level_1_data_device if self.obs_words == abi.OBS_L1 else self.env.level_2_data_device)(self.obs.ptr)
        return self.obs

    def reset(self) -> DeviceArray:
        self._n_steps = 0
        self.env.reset()   # (the agent population survives a reset; its state is cleared with the books)
        return self._observe()

    def step(self, actions) -> typing.Tuple[DeviceArray, DeviceArray]:
        nbytes = self.n_envs * self.rows * ACTION_DTYPE.itemsize
        ptr = _device_ptr(actions, nbytes)
        if ptr is None:  # host array: one H2D copy
            a = np.ascontiguousarray(actions, dtype=ACTION_DTYPE)
            if a.shape != (self.n_envs, self.rows):
                raise ValueError(f"actions must have shape ({self.n_envs}, {self.rows})")
            self._actions.copy_from_host(a)
            ptr = self._actions.ptr
        # an open-ended loop consumes each observation as it is produced: the library's per-step history is a scratch ring here
        self._n_steps += 1
        if self._n_steps > self.env.max_steps:
            self.env.clear_history()
            self._n_steps = 1
        # one launch: (background agents,) ids, Env::step and the observation records written straight into self.obs
        if self._agents:
            self.env.run_agents_with_rows(self._agent_seed, ptr, self._offsets.ptr, self.n_envs * self.rows, self.ids.ptr, self.obs.ptr)
        else:
            self.env.step_device(ptr, self._offsets.ptr, self.n_envs * self.rows, self.ids.ptr, self.obs.ptr)
        return self.obs, self.ids

    def check_errors(self):
        """Raise if any env flagged an error (capacity, bad order id, off-tick price); synchronises."""
        e = self.env.env_errors()
        if e.any():
            bad = np.flatnonzero(e)
            raise RuntimeError(f"{len(bad)} env(s) flagged errors, first env {bad[0]}: 0x{int(e[bad[0]]):x}")
