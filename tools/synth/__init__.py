"""ctypes binding of tools/synth/libexb_synth.so: the deterministic synthetic FASTA / FASTQ generators of SURVEY 8(d).

Test and bench tooling; the product library does not contain the generators."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libexb_synth.so")
GEN_FASTA, GEN_ILLUMINA, GEN_ONT = 1, 2, 4


class GenParams(C.Structure):
    _fields_ = [
        ("kind", C.c_int32),
        ("seed", C.c_uint64),
        ("n_records", C.c_int64),
        ("first_record", C.c_int64),
        ("len_min", C.c_int32),
        ("len_max", C.c_int32),
        ("wrap", C.c_int32),
        ("crlf", C.c_int32),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("%s is missing: run `make -C tools/synth` (or __graft_entry__.build())" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        L.exb_gen_size.restype = C.c_int64
        L.exb_gen_size.argtypes = [C.POINTER(GenParams)]
        L.exb_gen_size_device.restype = C.c_int
        L.exb_gen_size_device.argtypes = [C.POINTER(GenParams), C.POINTER(C.c_int64), C.c_void_p]
        L.exb_gen_device.restype = C.c_int
        L.exb_gen_device.argtypes = [C.POINTER(GenParams), C.c_void_p, C.c_int64, C.c_void_p]
        L.exb_gen_host.restype = C.c_int
        L.exb_gen_host.argtypes = [C.POINTER(GenParams), C.c_void_p, C.c_int64]
        _lib = L
    return _lib


def gen_params(kind, n_records, seed=1, first_record=0, len_min=150, len_max=150, wrap=60, crlf=False):
    p = GenParams()
    p.kind = {"fasta": GEN_FASTA, "illumina": GEN_ILLUMINA, "ont": GEN_ONT}[kind] if isinstance(kind, str) else kind
    p.seed = seed
    p.n_records = n_records
    p.first_record = first_record
    p.len_min = len_min
    p.len_max = len_max
    p.wrap = wrap
    p.crlf = 1 if crlf else 0
    return p


def gen_size(params):
    return lib().exb_gen_size(C.byref(params))


def gen_host(params):
    """The text as a numpy uint8 array (no GPU needed)."""
    import numpy as np

    size = gen_size(params)
    out = np.empty(size, dtype=np.uint8)
    rc = lib().exb_gen_host(C.byref(params), C.c_void_p(out.ctypes.data), size)
    if rc != 0:
        raise RuntimeError("exb_gen_host failed: %d" % rc)
    return out


def gen_device(params, device="cuda"):
    """The same text generated on the device: a torch uint8 tensor with 64 bytes of zeroed slack behind it."""
    import torch

    dev = torch.device(device)
    if params.n_records > 2_000_000:  # the host loop takes ~0.1 us per record: count on the device
        sz = C.c_int64()
        with torch.cuda.device(dev):
            rc = lib().exb_gen_size_device(C.byref(params), C.byref(sz), C.c_void_p(torch.cuda.current_stream().cuda_stream))
        if rc != 0:
            raise RuntimeError("exb_gen_size_device failed: %d" % rc)
        size = sz.value
    else:
        size = gen_size(params)
    buf = torch.empty(size + 64, dtype=torch.uint8, device=dev)
    buf[size:].zero_()
    with torch.cuda.device(buf.device):
        rc = lib().exb_gen_device(C.byref(params), C.c_void_p(buf.data_ptr()), size + 16, C.c_void_p(torch.cuda.current_stream().cuda_stream))
    if rc != 0:
        raise RuntimeError("exb_gen_device failed: %d" % rc)
    return buf[:size]
