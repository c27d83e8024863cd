"""Python binding over the C ABI -- the ctypes counterpart of the reference's cython module
``pylibwholegraph.binding.wholememory_binding`` (``wholememory_binding.pyx``): same class and
function names, same argument meaning, same exceptions per error code (pyx:255-277), so the
torch-level layer (``wholegraph_b200.torch``) reads like ``pylibwholegraph.torch``.

Views of WholeMemory are handed out as ``torch`` tensors built on the raw device/host pointer
(the reference exports DLPack capsules, pyx:1105-1295; zero-copy either way).
"""
import ctypes
import enum
from ctypes import POINTER, byref, c_int, c_int64, c_size_t, c_void_p

from . import _lib
from ._lib import lib


class WholeMemoryErrorCode(enum.IntEnum):
    Success = 0
    UnknowError = 1
    NotImplemented = 2
    LogicError = 3
    CUDAError = 4
    CommunicationError = 5
    InvalidInput = 6
    InvalidValue = 7
    OutOfMemory = 8
    NotSupported = 9
    SystemError = 10


class WholeMemoryMemoryType(enum.IntEnum):
    MtNone = 0
    MtContinuous = 1
    MtChunked = 2
    MtDistributed = 3
    MtHierarchy = 4


class WholeMemoryMemoryLocation(enum.IntEnum):
    MlNone = 0
    MlDevice = 1
    MlHost = 2


class WholeMemoryDistributedBackend(enum.IntEnum):
    DbNone = 0
    DbNCCL = 1
    DbNVSHMEM = 2


class WholeMemoryLogLevel(enum.IntEnum):
    LevFatal = 0
    LevError = 1
    LevWarn = 2
    LevInfo = 3
    LevDebug = 4
    LevTrace = 5


class WholeMemoryMemoryAllocType(enum.IntEnum):
    MatNone = 0
    MatDevice = 1
    MatHost = 2
    MatPinned = 3


class WholeMemoryDataType(enum.IntEnum):
    DtUnknown = 0
    DtFloat = 1
    DtHalf = 2
    DtDouble = 3
    DtBF16 = 4
    DtInt = 5
    DtInt64 = 6
    DtInt16 = 7
    DtInt8 = 8
    DtCount = 9


class WholeMemoryAccessType(enum.IntEnum):
    AtNone = 0
    AtReadOnly = 1
    AtReadWrite = 2


class WholeMemoryOptimizerType(enum.IntEnum):
    OptNone = 0
    OptSgd = 1
    OptLazyAdam = 2
    OptRmsProp = 3
    OptAdaGrad = 4


class WholeMemoryViewType(enum.IntEnum):
    VtNone = 0
    VtLocal = 1
    VtGlobal = 2
    VtRemote = 3


# export enum members at module level like a cpdef enum does
for _e in (WholeMemoryErrorCode, WholeMemoryMemoryType, WholeMemoryMemoryLocation, WholeMemoryDistributedBackend,
           WholeMemoryLogLevel, WholeMemoryMemoryAllocType, WholeMemoryDataType, WholeMemoryAccessType,
           WholeMemoryOptimizerType, WholeMemoryViewType):
    for _m in _e:
        if _m.name != "SystemError":  # not a member of the reference's enum; keep the builtin exception type visible
            globals()[_m.name] = _m   # incl. `NotImplemented` = 2, which the reference module exports as well


def check_wholememory_error_code(err):
    """Same mapping as the reference binding (pyx:255-277)."""
    err = int(err)
    if err == 0:
        return
    if err == WholeMemoryErrorCode.UnknowError:
        raise Exception("Unknown error")
    if err == WholeMemoryErrorCode.NotImplemented:
        raise NotImplementedError("Not implemented")
    if err == WholeMemoryErrorCode.LogicError:
        raise RuntimeError("Logic error")
    if err == WholeMemoryErrorCode.CUDAError:
        raise RuntimeError("CUDA error")
    if err == WholeMemoryErrorCode.CommunicationError:
        raise RuntimeError("Communication error")
    if err == WholeMemoryErrorCode.InvalidInput:
        raise ValueError("Invalid input")
    if err == WholeMemoryErrorCode.InvalidValue:
        raise ValueError("Invalid value")
    if err == WholeMemoryErrorCode.OutOfMemory:
        raise MemoryError("Out of memory")
    raise NotImplementedError("Error code %d not recognized" % err)


_chk = check_wholememory_error_code


def get_type_string(data_type):
    return {
        WholeMemoryDataType.DtFloat: "<f4", WholeMemoryDataType.DtHalf: "<f2", WholeMemoryDataType.DtDouble: "<f8",
        WholeMemoryDataType.DtBF16: "<f2", WholeMemoryDataType.DtInt: "<i4", WholeMemoryDataType.DtInt64: "<i8",
        WholeMemoryDataType.DtInt16: "<i2", WholeMemoryDataType.DtInt8: "|i1",
    }[WholeMemoryDataType(data_type)]


# --------------------------------------------------------------------------- init / ids
def init(flags=0, log_level=WholeMemoryLogLevel.LevInfo):
    _chk(lib.wholememory_init(flags, int(log_level)))


def finalize():
    _chk(lib.wholememory_finalize())


def fork_get_gpu_count():
    return lib.fork_get_device_count()


def py_get_wholememory_tensor_count():
    return lib.get_wholememory_tensor_count()


class PyWholeMemoryUniqueID:
    """128-byte communicator id; exposes a writable buffer so it can be broadcast."""

    def __init__(self):
        self.c = _lib.UniqueId()

    def __len__(self):
        return _lib.WHOLEMEMORY_UNIQUE_ID_BYTES

    def as_bytes(self):
        return ctypes.string_at(ctypes.addressof(self.c), len(self))

    def set_bytes(self, b):
        assert len(b) == len(self)
        ctypes.memmove(ctypes.addressof(self.c), bytes(b), len(self))

    def as_tensor(self):
        """int8 torch view sharing memory with the id (reference: uid.__dlpack__())."""
        import torch
        buf = (ctypes.c_int8 * len(self)).from_address(ctypes.addressof(self.c))
        t = torch.frombuffer(buf, dtype=torch.int8)
        t._wm_keepalive = self
        return t

    # The reference class exports its 128 bytes as an int8 [128] host DLPack tensor and through the buffer protocol
    # (pyx:1022-1067); pylibwholegraph/torch/comm.py broadcasts the id through `from_dlpack(uid.__dlpack__())`.
    def __dlpack__(self, stream=None):
        return self.as_tensor().__dlpack__()

    def __dlpack_device__(self):
        return (1, 0)  # kDLCPU

    def __buffer__(self, flags):
        return memoryview((ctypes.c_int8 * len(self)).from_address(ctypes.addressof(self.c)))


def create_unique_id():
    uid = PyWholeMemoryUniqueID()
    _chk(lib.wholememory_create_unique_id(byref(uid.c)))
    return uid


# --------------------------------------------------------------------------- communicator
class PyWholeMemoryComm:
    def __init__(self, c_handle=None):
        self.comm_id = c_void_p(c_handle)

    def get_c_handle(self):
        return self.comm_id.value

    def support_type_location(self, memory_type, memory_location):
        return lib.wholememory_communicator_support_type_location(self.comm_id, int(memory_type),
                                                                  int(memory_location)) == 0

    def get_rank(self):
        r = c_int(-1)
        _chk(lib.wholememory_communicator_get_rank(byref(r), self.comm_id))
        return r.value

    def get_size(self):
        s = c_int(-1)
        _chk(lib.wholememory_communicator_get_size(byref(s), self.comm_id))
        return s.value

    def get_clique_info(self):
        ci = _lib.CliqueInfo()
        _chk(lib.wholememory_communicator_get_clique_info(byref(ci), self.comm_id))
        cf = ci.clique_first_rank if ci.is_in_clique > 0 else -1
        cr = ci.clique_rank if ci.is_in_clique > 0 else -1
        cn = ci.clique_rank_num if ci.is_in_clique > 0 else -1
        return ci.is_in_clique > 0, cf, cr, cn, ci.clique_id, ci.clique_num

    def barrier(self):
        _chk(lib.wholememory_communicator_barrier(self.comm_id))

    def get_distributed_backend(self):
        return WholeMemoryDistributedBackend(lib.wholememory_communicator_get_distributed_backend(self.comm_id))

    def set_distributed_backend(self, distributed_backend):
        _chk(lib.wholememory_communicator_set_distributed_backend(self.comm_id, int(distributed_backend)))


def create_communicator(py_uid, world_rank, world_size):
    comm = c_void_p()
    _chk(lib.wholememory_create_communicator(byref(comm), py_uid.c, world_rank, world_size))
    return PyWholeMemoryComm(comm.value)


def destroy_communicator(py_comm):
    _chk(lib.wholememory_destroy_communicator(py_comm.comm_id))
    py_comm.comm_id = c_void_p(None)


def split_communicator(comm, color, key):
    new_comm = c_void_p()
    _chk(lib.wholememory_split_communicator(byref(new_comm), comm.comm_id, color, key))
    return PyWholeMemoryComm(new_comm.value)


def communicator_set_distributed_backend(py_comm, distributed_backend):
    py_comm.set_distributed_backend(distributed_backend)


def equal_partition_plan(entry_count, world_size):
    per = c_size_t(0)
    _chk(lib.wholememory_equal_entry_partition_plan(byref(per), entry_count, world_size))
    return per.value


# --------------------------------------------------------------------------- raw views as torch tensors
_TORCH_DTYPES = None


def _torch_dtype(dt):
    global _TORCH_DTYPES
    import torch
    if _TORCH_DTYPES is None:
        _TORCH_DTYPES = {
            WholeMemoryDataType.DtFloat: torch.float32, WholeMemoryDataType.DtHalf: torch.float16,
            WholeMemoryDataType.DtDouble: torch.float64, WholeMemoryDataType.DtBF16: torch.bfloat16,
            WholeMemoryDataType.DtInt: torch.int32, WholeMemoryDataType.DtInt64: torch.int64,
            WholeMemoryDataType.DtInt16: torch.int16, WholeMemoryDataType.DtInt8: torch.int8,
        }
    return _TORCH_DTYPES[WholeMemoryDataType(dt)]


class _CudaArray:
    """Minimal __cuda_array_interface__ provider over a raw device pointer."""

    def __init__(self, ptr, nbytes, owner):
        self.__cuda_array_interface__ = {
            "shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3, "strides": None,
        }
        self._owner = owner


def _view_as_torch(ptr, nbytes, dtype, location, device_id, owner):
    """uint8 view of [ptr, ptr+nbytes) reinterpreted as dtype.  location: MlDevice -> cuda tensor."""
    import torch
    tdt = _torch_dtype(dtype)
    if nbytes == 0 or not ptr:
        dev = ("cuda:%d" % device_id) if location == WholeMemoryMemoryLocation.MlDevice else "cpu"
        return torch.empty(0, dtype=tdt, device=dev)
    if location == WholeMemoryMemoryLocation.MlDevice:
        t = torch.as_tensor(_CudaArray(ptr, nbytes, owner), device="cuda:%d" % device_id)
    else:
        buf = (ctypes.c_uint8 * nbytes).from_address(ptr)
        t = torch.frombuffer(buf, dtype=torch.uint8)
        t._wm_keepalive = owner
    return t.view(tdt)


# --------------------------------------------------------------------------- handles / tensors
class PyWholeMemoryHandle:
    def __init__(self, c_handle=None):
        self.wholememory_handle = c_void_p(c_handle)

    def get_c_handle(self):
        return self.wholememory_handle.value

    def get_communicator(self):
        comm = c_void_p()
        _chk(lib.wholememory_get_communicator(byref(comm), self.wholememory_handle))
        return PyWholeMemoryComm(comm.value)

    def get_local_communicator(self):
        """HIERARCHY memory only: this build answers WHOLEMEMORY_NOT_SUPPORTED (NotImplementedError)."""
        comm = c_void_p()
        _chk(lib.wholememory_get_local_communicator(byref(comm), self.wholememory_handle))
        return PyWholeMemoryComm(comm.value)

    def get_cross_communicator(self):
        """HIERARCHY memory only: this build answers WHOLEMEMORY_NOT_SUPPORTED (NotImplementedError)."""
        comm = c_void_p()
        _chk(lib.wholememory_get_cross_communicator(byref(comm), self.wholememory_handle))
        return PyWholeMemoryComm(comm.value)

    def get_memory_type(self):
        return WholeMemoryMemoryType(lib.wholememory_get_memory_type(self.wholememory_handle))

    def get_memory_location(self):
        return WholeMemoryMemoryLocation(lib.wholememory_get_memory_location(self.wholememory_handle))

    def get_total_size(self):
        return lib.wholememory_get_total_size(self.wholememory_handle)

    def get_local_memory(self):
        p, s, o = c_void_p(), c_size_t(), c_size_t()
        _chk(lib.wholememory_get_local_memory(byref(p), byref(s), byref(o), self.wholememory_handle))
        return p.value, s.value, o.value

    def get_rank_memory(self, rank):
        p, s, o = c_void_p(), c_size_t(), c_size_t()
        _chk(lib.wholememory_get_rank_memory(byref(p), byref(s), byref(o), rank, self.wholememory_handle))
        return p.value, s.value, o.value

    def get_global_pointer(self):
        p = c_void_p()
        _chk(lib.wholememory_get_global_pointer(byref(p), self.wholememory_handle))
        return p.value

    def get_global_reference(self):
        g = _lib.GlobalReference()
        _chk(lib.wholememory_get_global_reference(byref(g), self.wholememory_handle))
        return g

    def get_rank_partition_offsets(self):
        n = self.get_communicator().get_size()
        arr = (c_size_t * (n + 1))()
        _chk(lib.wholememory_get_rank_partition_offsets(arr, self.wholememory_handle))
        return list(arr)

    # torch views (reference: get_local_flatten_tensor / get_global_flatten_tensor / get_all_chunked_flatten_tensor).
    # Both call forms are accepted: this binding's (dtype, location, device id) and the reference's, whose first argument
    # is the DLPack importer (pyx:1368-1412); the importer is then applied to the view (torch tensors speak __dlpack__).
    def get_local_flatten_tensor(self, *args):
        fn, (dtype, view_from_location, view_from_device_id) = _split_importer(args)
        t, off = self._local_flat(dtype, view_from_location, view_from_device_id)
        return (fn(t) if fn else t), off

    def get_global_flatten_tensor(self, *args):
        fn, (dtype, view_from_location, view_from_device_id) = _split_importer(args)
        t, off = self._global_flat(dtype, view_from_location, view_from_device_id)
        return (fn(t) if fn else t), off

    def get_all_chunked_flatten_tensor(self, *args):
        fn, (dtype, view_from_location, view_from_device_id) = _split_importer(args)
        ts, offs = self._chunked_flat(dtype, view_from_location, view_from_device_id)
        return ([fn(t) for t in ts] if fn else ts), offs

    def _local_flat(self, dtype, view_from_location, view_from_device_id):
        p, s, o = self.get_local_memory()
        es = lib.wholememory_dtype_get_element_size(int(dtype))
        return _view_as_torch(p, s, dtype, self._view_loc(view_from_location), view_from_device_id, self), o // es

    def _global_flat(self, dtype, view_from_location, view_from_device_id):
        p = self.get_global_pointer()
        return _view_as_torch(p, self.get_total_size(), dtype, self._view_loc(view_from_location),
                              view_from_device_id, self), 0

    def _chunked_flat(self, dtype, view_from_location, view_from_device_id):
        n = self.get_communicator().get_size()
        es = lib.wholememory_dtype_get_element_size(int(dtype))
        tensors, offsets = [], []
        for r in range(n):
            p, s, o = self.get_rank_memory(r)
            tensors.append(_view_as_torch(p, s, dtype, self._view_loc(view_from_location), view_from_device_id, self))
            offsets.append(o // es)
        return tensors, offsets

    def _view_loc(self, view_from_location):
        loc = self.get_memory_location()
        view = WholeMemoryMemoryLocation(view_from_location)
        if loc == WholeMemoryMemoryLocation.MlDevice and view == WholeMemoryMemoryLocation.MlHost:
            raise ValueError("Device WholeMemory cannot get view from host.")
        return view

    def from_filelist(self, memory_offset, memory_entry_size, file_entry_size, round_robin_size, file_list):
        load_wholememory_handle_from_filelist(self.wholememory_handle.value, memory_offset, memory_entry_size, file_entry_size,
                                              round_robin_size, file_list)

    def to_file(self, memory_offset, memory_entry_size, file_entry_size, file_name):
        store_wholememory_handle_to_file(self.wholememory_handle.value, memory_offset, memory_entry_size, file_entry_size, file_name)


def _split_importer(args):
    """(import_dlpack_fn or None, the remaining three view arguments)"""
    if len(args) == 4 and callable(args[0]):
        return args[0], args[1:]
    if len(args) != 3:
        raise TypeError("expected ([import_dlpack_fn,] dtype, view_from_location, view_from_device_id)")
    return None, args


def load_wholememory_handle_from_filelist(wholememory_handle_int_ptr, memory_offset, memory_entry_size, file_entry_size,
                                          round_robin_size, file_list):
    """Every rank loads its partition from the list of raw binary files (reference pyx:1836-1861)."""
    arr = (ctypes.c_char_p * len(file_list))(*[f.encode() for f in file_list])
    _chk(lib.wholememory_load_from_file(c_void_p(wholememory_handle_int_ptr), memory_offset, memory_entry_size, file_entry_size,
                                        arr, len(file_list), round_robin_size))


def store_wholememory_handle_to_file(wholememory_handle_int_ptr, memory_offset, memory_entry_size, file_entry_size, file_name):
    """Every rank stores its partition to its own file (reference pyx:1863-1873)."""
    _chk(lib.wholememory_store_to_file(c_void_p(wholememory_handle_int_ptr), memory_offset, memory_entry_size, file_entry_size,
                                       file_name.encode()))


class PyWholeMemoryTensorDescription:
    def __init__(self):
        # all-zero like the reference class (pyx:1438-1441): dim 0, dtype unknown, sizes / strides unset until the setters run
        self.tensor_description = _lib.TensorDescription()

    def set_dtype(self, dtype):
        self.tensor_description.dtype = int(dtype)

    def set_shape(self, shape):
        assert 0 < len(shape) < _lib.WHOLEMEMORY_MAX_TENSOR_DIM  # 1..7 dims, like the reference class (pyx:1450)
        self.tensor_description.dim = len(shape)
        for i, s in enumerate(shape):
            self.tensor_description.sizes[i] = int(s)

    def set_stride(self, strides):
        assert len(strides) == self.tensor_description.dim
        for i, s in enumerate(strides):
            self.tensor_description.strides[i] = int(s)

    def set_storage_offset(self, storage_offset):
        self.tensor_description.storage_offset = int(storage_offset)

    @property
    def dtype(self):
        return WholeMemoryDataType(self.tensor_description.dtype)

    def dim(self):
        return self.tensor_description.dim

    @property
    def shape(self):
        return tuple(self.tensor_description.sizes[i] for i in range(self.dim()))

    def stride(self):
        return tuple(self.tensor_description.strides[i] for i in range(self.dim()))

    def storage_offset(self):
        return self.tensor_description.storage_offset


class WrappedLocalTensor:
    """Non-owning wholememory_tensor_t over caller memory (indices, outputs, gradients)."""

    def __init__(self):
        self.wm_tensor = c_void_p(None)

    def __del__(self):
        # (module globals are already cleared when this runs at interpreter shutdown: nothing left to free then)
        if lib is not None and getattr(self, "wm_tensor", None) is not None and self.wm_tensor.value:
            lib.wholememory_destroy_tensor(self.wm_tensor)
            self.wm_tensor = c_void_p(None)

    def wrap_tensor(self, py_desc, data_ptr):
        _chk(lib.wholememory_make_tensor_from_pointer(byref(self.wm_tensor), c_void_p(data_ptr),
                                                      byref(py_desc.tensor_description)))
        return self

    def get_c_handle(self):
        return self.wm_tensor.value or 0


class PyWholeMemoryTensor:
    def __init__(self, c_tensor=None):
        self.wholememory_tensor = c_void_p(c_tensor)

    def get_c_handle(self):
        return self.wholememory_tensor.value

    def _desc(self):
        return lib.wholememory_tensor_get_tensor_description(self.wholememory_tensor).contents

    def get_wholememory_handle(self):
        return PyWholeMemoryHandle(lib.wholememory_tensor_get_memory_handle(self.wholememory_tensor))

    @property
    def dtype(self):
        return WholeMemoryDataType(self._desc().dtype)

    def dim(self):
        return self._desc().dim

    @property
    def shape(self):
        d = self._desc()
        return tuple(d.sizes[i] for i in range(d.dim))

    def stride(self):
        d = self._desc()
        return tuple(d.strides[i] for i in range(d.dim))

    def storage_offset(self):
        return self._desc().storage_offset

    def get_local_entry_count(self):
        v = c_size_t()
        _chk(lib.wholememory_tensor_get_local_entry_count(byref(v), self.wholememory_tensor))
        return v.value

    def get_local_entry_start(self):
        v = c_size_t()
        _chk(lib.wholememory_tensor_get_local_entry_start(byref(v), self.wholememory_tensor))
        return v.value

    def get_entry_offsets(self):
        n = self.get_wholememory_handle().get_communicator().get_size()
        arr = (c_size_t * (n + 1))()
        _chk(lib.wholememory_tensor_get_entry_offsets(arr, self.wholememory_tensor))
        return list(arr)

    def get_sub_tensor(self, starts, ends):
        d = self.dim()
        if len(starts) != d or len(ends) != d:
            raise ValueError("starts/ends must have one entry per dim")
        s = (c_int64 * d)(*starts)
        e = (c_int64 * d)(*ends)
        sub = c_void_p()
        _chk(lib.wholememory_tensor_get_subtensor(self.wholememory_tensor, s, e, byref(sub)))
        return PyWholeMemoryTensor(sub.value)

    def get_tensor_in_window(self, flatten_tensor, storage_window_offset):
        """Reshape a flat view (element offset `storage_window_offset` from the handle start) to this tensor's
        window, same arithmetic as the reference (pyx:1584-1610)."""
        d = self._desc()
        if d.dim == 1:
            start_indice = max(0, d.storage_offset - storage_window_offset)
            end_indice = min(flatten_tensor.shape[0], d.storage_offset + d.sizes[0] - storage_window_offset)
            return flatten_tensor[start_indice:end_indice], max(0, storage_window_offset - d.storage_offset)
        embedding_stride = d.strides[0]
        storage_offset0 = d.storage_offset // embedding_stride
        storage_offset1 = d.storage_offset % embedding_stride
        mat = flatten_tensor.reshape(-1, embedding_stride)
        assert storage_window_offset % embedding_stride == 0
        vector_start_offset = storage_window_offset // embedding_stride
        start_indice0 = max(0, storage_offset0 - vector_start_offset)
        end_indice0 = min(mat.shape[0], storage_offset0 + d.sizes[0] - vector_start_offset)
        return (mat[start_indice0:end_indice0, storage_offset1:storage_offset1 + d.sizes[1]],
                max(0, vector_start_offset - storage_offset0))

    # like the handle's getters, these also take the reference's form with the DLPack importer first (pyx:1612-1650)
    def get_local_tensor(self, *args):
        fn, (view_from_location, view_from_device_id) = _split_importer2(args)
        flat, off = self.get_wholememory_handle().get_local_flatten_tensor(*_with_importer(fn, self.dtype, view_from_location,
                                                                                           view_from_device_id))
        return self.get_tensor_in_window(flat, off)

    def get_global_tensor(self, *args):
        fn, (view_from_location, view_from_device_id) = _split_importer2(args)
        flat, _ = self.get_wholememory_handle().get_global_flatten_tensor(*_with_importer(fn, self.dtype, view_from_location,
                                                                                          view_from_device_id))
        return self.get_tensor_in_window(flat, 0)[0]

    def get_all_chunked_tensor(self, *args):
        fn, (view_from_location, view_from_device_id) = _split_importer2(args)
        ts, offs = self.get_wholememory_handle().get_all_chunked_flatten_tensor(*_with_importer(fn, self.dtype, view_from_location,
                                                                                                view_from_device_id))
        out_t, out_o = [], []
        for t, o in zip(ts, offs):
            tw, ow = self.get_tensor_in_window(t, o)
            out_t.append(tw)
            out_o.append(ow)
        return out_t, out_o

    def from_filelist(self, filelist, round_robin_size=0):
        d = self._desc()
        es = lib.wholememory_dtype_get_element_size(d.dtype)
        stride = d.strides[0] if d.dim == 2 else 1
        cols = d.sizes[1] if d.dim == 2 else 1
        self.get_wholememory_handle().from_filelist(d.storage_offset * es, stride * es, cols * es, round_robin_size,
                                                    filelist)

    def to_file(self, filename):
        d = self._desc()
        es = lib.wholememory_dtype_get_element_size(d.dtype)
        stride = d.strides[0] if d.dim == 2 else 1
        cols = d.sizes[1] if d.dim == 2 else 1
        self.get_wholememory_handle().to_file(d.storage_offset * es, stride * es, cols * es, filename)


def _split_importer2(args):
    if len(args) == 3 and callable(args[0]):
        return args[0], args[1:]
    if len(args) != 2:
        raise TypeError("expected ([import_dlpack_fn,] view_from_location, view_from_device_id)")
    return None, args


def _with_importer(fn, *rest):
    return ((fn,) + rest) if fn else rest


def malloc(total_size, py_comm, memory_type, memory_location, data_granularity, rank_entry_partition=None):
    h = c_void_p()
    part = None
    if rank_entry_partition is not None:
        part = (c_size_t * len(rank_entry_partition))(*rank_entry_partition)
    _chk(lib.wholememory_malloc(byref(h), total_size, py_comm.comm_id, int(memory_type), int(memory_location),
                                data_granularity, part))
    return PyWholeMemoryHandle(h.value)


def free(handle):
    _chk(lib.wholememory_free(handle.wholememory_handle))


def create_wholememory_tensor(tensor_description, comm, memory_type, memory_location, tensor_entry_partition=None):
    if tensor_description.dim() not in (1, 2):
        raise NotImplementedError("WholeMemory currently only support 1D or 2D tensor")
    if tensor_description.stride()[tensor_description.dim() - 1] != 1:
        raise ValueError("last stride should be 1")
    if tensor_description.storage_offset() != 0:
        raise ValueError("storage_offset be 0 when created")
    t = c_void_p()
    part = None
    if tensor_entry_partition is not None:
        part = (c_size_t * len(tensor_entry_partition))(*tensor_entry_partition)
    _chk(lib.wholememory_create_tensor(byref(t), byref(tensor_description.tensor_description), comm.comm_id,
                                       int(memory_type), int(memory_location), part))
    return PyWholeMemoryTensor(t.value)


def create_wholememory_array(dtype, size, comm, mem_type, mem_location, tensor_entry_partition=None):
    d = PyWholeMemoryTensorDescription()
    d.set_dtype(dtype)
    d.set_shape((size,))
    d.set_stride((1,))
    return create_wholememory_tensor(d, comm, mem_type, mem_location, tensor_entry_partition)


def create_wholememory_matrix(dtype, row, column, stride, comm, mem_type, mem_location, tensor_entry_partition=None):
    d = PyWholeMemoryTensorDescription()
    d.set_dtype(dtype)
    d.set_shape((row, column))
    d.set_stride((column if stride == -1 else stride, 1))
    return create_wholememory_tensor(d, comm, mem_type, mem_location, tensor_entry_partition)


def make_tensor_as_wholememory(tensor_description, data_ptr):
    t = c_void_p()
    _chk(lib.wholememory_make_tensor_from_pointer(byref(t), c_void_p(data_ptr),
                                                  byref(tensor_description.tensor_description)))
    return PyWholeMemoryTensor(t.value)


def make_handle_as_wholememory(tensor_description, handle):
    t = c_void_p()
    _chk(lib.wholememory_make_tensor_from_handle(byref(t), handle.wholememory_handle,
                                                 byref(tensor_description.tensor_description)))
    return PyWholeMemoryTensor(t.value)


def destroy_wholememory_tensor(wholememory_tensor):
    _chk(lib.wholememory_destroy_tensor(wholememory_tensor.wholememory_tensor))
    wholememory_tensor.wholememory_tensor = c_void_p(None)


# --------------------------------------------------------------------------- ops
def _env_ptr(p_env_fns_int):
    return ctypes.cast(c_void_p(p_env_fns_int), POINTER(_lib.EnvFns))


# --------------------------------------------------------------------------- allocation callbacks, reference protocol
class PyMemoryAllocType:
    """The allocation kind handed to a malloc callback (reference pyx:348-364)."""

    def __init__(self):
        self.alloc_type = int(WholeMemoryMemoryAllocType.MatNone)

    def set_type(self, new_type):
        self.alloc_type = int(new_type)

    def get_type(self):
        return self.alloc_type

    set_ctype = set_type
    get_ctype = get_type


class GlobalContextWrapper:
    """wholememory_env_func_t built from Python callables, with the reference binding's calling convention
    (pyx:366-552), so code written against pylibwholegraph's binding runs unchanged:

        temp_create_context_fn(global_context) -> memory_context (any Python object; kept alive until destroyed)
        temp_destroy_context_fn(memory_context, global_context)
        malloc_fn(PyWholeMemoryTensorDescription, PyMemoryAllocType, memory_context, global_context) -> address (int)
        free_fn(memory_context, global_context)

    OUTPUT memory contexts are Python objects owned by the caller, passed to the ops as id(obj).
    (wholegraph_b200.torch.wholegraph_env carries a leaner torch-specific table; this one is the general form.)"""

    def __init__(self):
        self.env = _lib.EnvFns()
        self._fns = None
        self._live_temp = {}  # id -> memory context created on behalf of the library

    def create_context(self, temp_create_context_fn, temp_destroy_context_fn, temp_malloc_fn, temp_free_fn, temp_global_context,
                       output_malloc_fn, output_free_fn, output_global_context):
        def obj(address):
            return ctypes.cast(c_void_p(address), ctypes.py_object).value

        def describe(desc_ptr, malloc_type):
            d = PyWholeMemoryTensorDescription()
            ctypes.memmove(byref(d.tensor_description), desc_ptr, ctypes.sizeof(_lib.TensorDescription))
            k = PyMemoryAllocType()
            k.set_type(malloc_type)
            return d, k

        def c_create(out_ctx, _g):
            ctx = temp_create_context_fn(temp_global_context)
            self._live_temp[id(ctx)] = ctx
            out_ctx[0] = id(ctx)

        def c_destroy(memory_context, _g):
            ctx = self._live_temp.pop(memory_context, None)
            if ctx is not None:
                temp_destroy_context_fn(ctx, temp_global_context)

        def c_temp_malloc(desc_ptr, malloc_type, memory_context, _g):
            d, k = describe(desc_ptr, malloc_type)
            return int(temp_malloc_fn(d, k, self._live_temp[memory_context], temp_global_context) or 0)

        def c_temp_free(memory_context, _g):
            ctx = self._live_temp.get(memory_context)
            if ctx is not None:
                temp_free_fn(ctx, temp_global_context)

        def c_out_malloc(desc_ptr, malloc_type, memory_context, _g):
            d, k = describe(desc_ptr, malloc_type)
            return int(output_malloc_fn(d, k, obj(memory_context), output_global_context) or 0)

        def c_out_free(memory_context, _g):
            output_free_fn(obj(memory_context), output_global_context)

        self._fns = (_lib.CREATE_CTX_FN(c_create), _lib.DESTROY_CTX_FN(c_destroy), _lib.MALLOC_FN(c_temp_malloc), _lib.FREE_FN(c_temp_free),
                     _lib.MALLOC_FN(c_out_malloc), _lib.FREE_FN(c_out_free))
        t, o = self.env.temporary_fns, self.env.output_fns
        t.create_memory_context_fn, t.destroy_memory_context_fn, t.malloc_fn, t.free_fn = self._fns[:4]
        o.malloc_fn, o.free_fn = self._fns[4:]
        t.global_context = o.global_context = None

    def get_env_fns(self) -> int:
        return ctypes.addressof(self.env)


def wholememory_env_test_cython_op(input, output, output_variable_device_tensor_handle, output_variable_pinned_tensor_handle,
                                   output_variable_host_tensor_handle, output_variable_entry_count, p_env_fns_int, stream_int):
    """Allocator-plumbing self test (reference pyx:1919-1935): output[i, :] = input + i, also into three variable-size
    outputs allocated through the env functions (device, pinned, host)."""
    _chk(lib.wholememory_env_test_op(c_void_p(input.get_c_handle()), c_void_p(output.get_c_handle()),
                                     c_void_p(output_variable_device_tensor_handle), c_void_p(output_variable_pinned_tensor_handle),
                                     c_void_p(output_variable_host_tensor_handle), output_variable_entry_count, _env_ptr(p_env_fns_int),
                                     c_void_p(stream_int)))


class DLDeviceType(enum.IntEnum):
    kDLCPU = 1
    kDLCUDA = 2
    kDLCUDAHost = 3


globals().update(DLDeviceType.__members__)  # module-level constants, like a cpdef enum


class PyWholeMemoryFlattenDlpack:
    """DLPack-exporting flat view of a WholeMemory handle (reference pyx:1105-1290).  The reference builds the DLPack
    capsule by hand; here the view is a torch tensor over the same memory, which already speaks the protocol."""

    typestr = None

    def __init__(self):
        self.device_type = WholeMemoryMemoryLocation.MlHost
        self.device_id = 0
        self._view = None

    @property
    def ptr(self):
        return self._view.data_ptr() if self._view is not None else 0

    def set_view_device(self, device_type, device_id):
        self.device_type, self.device_id = WholeMemoryMemoryLocation(device_type), int(device_id)

    def get_view(self, handle, data_type, view_type, target_rank):
        """-> (element count, element offset of the view inside the whole allocation)"""
        view_type = WholeMemoryViewType(view_type)
        if view_type == WholeMemoryViewType.VtLocal:
            self._view, off = handle._local_flat(data_type, self.device_type, self.device_id)
        elif view_type == WholeMemoryViewType.VtGlobal:
            self._view, off = handle._global_flat(data_type, self.device_type, self.device_id)
        else:
            views, offs = handle._chunked_flat(data_type, self.device_type, self.device_id)
            self._view, off = views[target_rank], offs[target_rank]
        self.typestr = get_type_string(data_type)
        return self._view.numel(), off

    def __len__(self):
        """element count of the current view (reference pyx:1215-1216)"""
        return int(self._view.numel()) if self._view is not None else 0

    def __dlpack__(self, stream=None):
        return self._view.__dlpack__() if stream is None else self._view.__dlpack__(stream=stream)

    def __dlpack_device__(self):
        return self._view.__dlpack_device__()


def wholememory_gather_op(wholememory_tensor, indices_tensor, output_tensor, p_env_fns_int, stream_int,
                          gather_sms=-1):
    _chk(lib.wholememory_gather(wholememory_tensor.wholememory_tensor, c_void_p(indices_tensor.get_c_handle()),
                                c_void_p(output_tensor.get_c_handle()), _env_ptr(p_env_fns_int),
                                c_void_p(stream_int), gather_sms))


def wholememory_scatter_op(input_tensor, indices_tensor, wholememory_tensor, p_env_fns_int, stream_int,
                           scatter_sms=-1):
    _chk(lib.wholememory_scatter(c_void_p(input_tensor.get_c_handle()), c_void_p(indices_tensor.get_c_handle()),
                                 wholememory_tensor.wholememory_tensor, _env_ptr(p_env_fns_int),
                                 c_void_p(stream_int), scatter_sms))


def csr_unweighted_sample_without_replacement(wm_csr_row_ptr_tensor, wm_csr_col_ptr_tensor, center_nodes_tensor,
                                              max_sample_count, output_sample_offset_tensor,
                                              output_dest_memory_handle, output_center_localid_memory_handle,
                                              output_edge_gid_memory_handle, random_seed, p_env_fns_int, stream_int):
    _chk(lib.wholegraph_csr_unweighted_sample_without_replacement(
        wm_csr_row_ptr_tensor.wholememory_tensor, wm_csr_col_ptr_tensor.wholememory_tensor,
        c_void_p(center_nodes_tensor.get_c_handle()), max_sample_count,
        c_void_p(output_sample_offset_tensor.get_c_handle()), c_void_p(output_dest_memory_handle),
        c_void_p(output_center_localid_memory_handle), c_void_p(output_edge_gid_memory_handle), random_seed,
        _env_ptr(p_env_fns_int), c_void_p(stream_int)))


def csr_weighted_sample_without_replacement(wm_csr_row_ptr_tensor, wm_csr_col_ptr_tensor, wm_csr_weight_ptr_tensor,
                                            center_nodes_tensor, max_sample_count, output_sample_offset_tensor,
                                            output_dest_memory_handle, output_center_localid_memory_handle,
                                            output_edge_gid_memory_handle, random_seed, p_env_fns_int, stream_int):
    _chk(lib.wholegraph_csr_weighted_sample_without_replacement(
        wm_csr_row_ptr_tensor.wholememory_tensor, wm_csr_col_ptr_tensor.wholememory_tensor,
        wm_csr_weight_ptr_tensor.wholememory_tensor, c_void_p(center_nodes_tensor.get_c_handle()), max_sample_count,
        c_void_p(output_sample_offset_tensor.get_c_handle()), c_void_p(output_dest_memory_handle),
        c_void_p(output_center_localid_memory_handle), c_void_p(output_edge_gid_memory_handle), random_seed,
        _env_ptr(p_env_fns_int), c_void_p(stream_int)))


def host_generate_exponential_distribution_negative_float(random_seed, sub_sequence, output):
    _chk(lib.generate_exponential_distribution_negative_float_cpu(random_seed, sub_sequence, c_void_p(output.get_c_handle())))


def append_unique(target_node_tensor, neighbor_node_tensor, output_unique_node_memory_handle,
                  output_neighbor_raw_to_unique_mapping_tensor, p_env_fns_int, stream_int):
    _chk(lib.graph_append_unique(c_void_p(target_node_tensor.get_c_handle()), c_void_p(neighbor_node_tensor.get_c_handle()),
                                 c_void_p(output_unique_node_memory_handle),
                                 c_void_p(output_neighbor_raw_to_unique_mapping_tensor.get_c_handle()),
                                 _env_ptr(p_env_fns_int), c_void_p(stream_int)))


def add_csr_self_loop(csr_row_ptr_tensor, csr_col_ptr_tensor, csr_row_ptr_self_tensor, csr_col_ptr_self_tensor, stream_int):
    _chk(lib.csr_add_self_loop(c_void_p(csr_row_ptr_tensor.get_c_handle()), c_void_p(csr_col_ptr_tensor.get_c_handle()),
                               c_void_p(csr_row_ptr_self_tensor.get_c_handle()), c_void_p(csr_col_ptr_self_tensor.get_c_handle()),
                               c_void_p(stream_int)))


def host_generate_random_positive_int(random_seed, sub_sequence, output):
    _chk(lib.generate_random_positive_int_cpu(random_seed, sub_sequence, c_void_p(output.get_c_handle())))


# --------------------------------------------------------------------------- embedding
class WholeMemoryOptimizer:
    param_dict = None

    def __init__(self):
        self.wm_optimizer = c_void_p(None)
        self.optimizer_type = WholeMemoryOptimizerType.OptNone
        self.param_dict = None

    def create_optimizer(self, optimizer_type, param_dict):
        self.optimizer_type = WholeMemoryOptimizerType(optimizer_type)
        self.param_dict = param_dict
        _chk(lib.wholememory_create_embedding_optimizer(byref(self.wm_optimizer), int(optimizer_type)))
        for key, value in param_dict.items():
            f = ctypes.c_float(float(value))
            _chk(lib.wholememory_optimizer_set_parameter(self.wm_optimizer, key.encode("utf-8"),
                                                         ctypes.cast(byref(f), c_void_p)))

    def add_embedding(self, embedding):
        _chk(lib.wholememory_embedding_set_optimizer(embedding.wm_embedding, self.wm_optimizer))

    def destroy_optimizer(self):
        if not self.wm_optimizer.value:
            return
        lib.wholememory_destroy_embedding_optimizer(self.wm_optimizer)
        self.wm_optimizer = c_void_p(None)
        self.optimizer_type = WholeMemoryOptimizerType.OptNone
        self.param_dict = None


def create_optimizer(optimizer_type, param_dict):
    o = WholeMemoryOptimizer()
    o.create_optimizer(optimizer_type, param_dict)
    return o


def create_non_optimizer():
    return WholeMemoryOptimizer()


class WholeMemoryCachePolicy:
    def __init__(self):
        self.cache_policy = c_void_p(None)

    def create_policy(self, comm, memory_type, memory_location, access_type, ratio):
        _chk(lib.wholememory_create_embedding_cache_policy(byref(self.cache_policy), comm.comm_id, int(memory_type),
                                                           int(memory_location), int(access_type), ratio))

    def destroy_policy(self):
        if not self.cache_policy.value:
            return
        _chk(lib.wholememory_destroy_embedding_cache_policy(self.cache_policy))
        self.cache_policy = c_void_p(None)


def create_cache_policy(comm, memory_type, memory_location, access_type, ratio):
    p = WholeMemoryCachePolicy()
    p.create_policy(comm, memory_type, memory_location, access_type, ratio)
    return p


def create_non_cache_policy():
    return WholeMemoryCachePolicy()


class PyWholeMemoryEmbedding:
    def __init__(self):
        self.wm_embedding = c_void_p(None)

    def create_embedding(self, tensor_desc, comm, memory_type, memory_location, cache_policy, embedding_entry_partition,
                         user_defined_sms, round_robin_size):
        part = None
        if embedding_entry_partition is not None:
            part = (c_size_t * len(embedding_entry_partition))(*embedding_entry_partition)
        _chk(lib.wholememory_create_embedding(byref(self.wm_embedding), byref(tensor_desc.tensor_description),
                                              comm.comm_id, int(memory_type), int(memory_location),
                                              cache_policy.cache_policy, part, user_defined_sms, round_robin_size))

    def destroy_embedding(self):
        _chk(lib.wholememory_destroy_embedding(self.wm_embedding))
        self.wm_embedding = c_void_p(None)

    def writeback_all_cache(self, stream):
        _chk(lib.wholememory_embedding_writeback_cache(self.wm_embedding, stream))

    def drop_all_cache(self, stream):
        _chk(lib.wholememory_embedding_drop_all_cache(self.wm_embedding, stream))

    def get_embedding_tensor(self):
        return PyWholeMemoryTensor(lib.wholememory_embedding_get_embedding_tensor(self.wm_embedding))

    def get_optimizer_state_names(self):
        names = lib.wholememory_embedding_get_optimizer_state_names(self.wm_embedding)
        out, i = [], 0
        while names and names[i] is not None:
            out.append(names[i].decode("utf-8"))
            i += 1
        return out

    def get_optimizer_state(self, state_name):
        return PyWholeMemoryTensor(
            lib.wholememory_embedding_get_optimizer_state(self.wm_embedding, state_name.encode("utf-8")))


def create_embedding(tensor_desc, comm, memory_type, memory_location, cache_policy, embedding_entry_partition=None,
                     user_defined_sms=-1, round_robin_size=0):
    e = PyWholeMemoryEmbedding()
    e.create_embedding(tensor_desc, comm, memory_type, memory_location, cache_policy, embedding_entry_partition,
                       user_defined_sms, round_robin_size)
    return e


def EmbeddingGatherForward(wm_embedding, indice, output, adjust_cache, p_env_fns_int, stream_int):
    _chk(lib.wholememory_embedding_gather(wm_embedding.wm_embedding, c_void_p(indice.get_c_handle()),
                                          c_void_p(output.get_c_handle()), adjust_cache, _env_ptr(p_env_fns_int),
                                          stream_int))


def EmbeddingGatherGradientApply(wm_embedding, indice, grads, adjust_cache, lr, p_env_fns_int, stream_int):
    _chk(lib.wholememory_embedding_gather_gradient_apply(wm_embedding.wm_embedding, c_void_p(indice.get_c_handle()),
                                                         c_void_p(grads.get_c_handle()), adjust_cache, lr,
                                                         _env_ptr(p_env_fns_int), stream_int))
