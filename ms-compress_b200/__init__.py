"""mscompress_b200 -- Python binding of the B200-native libMSCompression (test / benchmark convenience).

The product is the C-ABI shared library `lib/libMSCompression.so` (CUDA kernels for sm_100a + the reference's
`ms_compress` / `ms_decompress` / per-format one-shot functions, see include/*.h).  This module only binds it:

* `OpenSrc`            -- the same class the reference's Python harness exposes (test/compressors.py:214-314):
                          `OpenSrc(format).Compress(data)`, `.Decompress(data)`, `.MaxCompressedSize(n)`, host bytes in / out.
* `GpuContext`         -- the device-resident batch interface (include/mscomp_gpu.h) over torch CUDA tensors; torch
                          supplies device memory and the stream, nothing else.
* `generate(...)`      -- the synthetic byte streams of SURVEY.md section 8(d).

There is no CPU fallback: if the library has not been built, loading raises.
"""
import ctypes
import os
import struct

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_DIR = os.path.join(_HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libMSCompression.so")
GEN_PATH = os.path.join(LIB_DIR, "libmsdatagen.so")

MSCOMP_OK, MSCOMP_ARG_ERROR, MSCOMP_DATA_ERROR, MSCOMP_MEM_ERROR, MSCOMP_BUF_ERROR = 0, -2, -3, -4, -5


class CompressionFormat:
    """MSCompFormat (include/mscomp/general.h); names follow the reference's test/compressors.py."""
    NoCompression = 0
    LZNT1 = 2
    Xpress = 3
    XpressHuffman = 4


FORMAT_NAMES = {CompressionFormat.LZNT1: "lznt1", CompressionFormat.Xpress: "xpress", CompressionFormat.XpressHuffman: "xpress_huff"}

_lib = None
_gen = None


class MSCompError(Exception):
    def __init__(self, func, status):
        super().__init__("%s failed: %d" % (func, status))
        self.status = status


def build(force=False, verbose=False):
    """Compile the native libraries in-tree (nvcc, sm_100a)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("_mscomp_build", os.path.join(_HERE, "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.build(force=force, verbose=verbose)


def lib():
    """The loaded libMSCompression.so (ctypes).  Raises if it has not been built -- there is no other code path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("mscompress_b200: %s is missing -- run `python __graft_entry__.py build` (nvcc); "
                              "this package has no CPU or PyTorch fallback" % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        vp, sz, i32, u32, u64 = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_uint32, ctypes.c_uint64
        psz = ctypes.POINTER(ctypes.c_size_t)
        for name in ("lznt1", "xpress", "xpress_huff"):
            if not hasattr(L, name + "_compress"):
                continue
            for op in ("compress", "decompress"):
                f = getattr(L, "%s_%s" % (name, op))
                f.restype, f.argtypes = i32, [vp, sz, vp, psz]
            f = getattr(L, name + "_max_compressed_size")
            f.restype, f.argtypes = sz, [sz]
        L.ms_compress.restype, L.ms_compress.argtypes = i32, [i32, vp, sz, vp, psz]
        L.ms_decompress.restype, L.ms_decompress.argtypes = i32, [i32, vp, sz, vp, psz]
        L.ms_max_compressed_size.restype, L.ms_max_compressed_size.argtypes = sz, [i32, sz]
        L.mscomp_gpu_create.restype, L.mscomp_gpu_create.argtypes = vp, [i32]
        L.mscomp_gpu_create_on_stream.restype, L.mscomp_gpu_create_on_stream.argtypes = vp, [i32, vp]
        L.mscomp_gpu_destroy.restype, L.mscomp_gpu_destroy.argtypes = None, [vp]
        L.mscomp_gpu_synchronize.restype, L.mscomp_gpu_synchronize.argtypes = i32, [vp]
        L.mscomp_gpu_kernel_launches.restype, L.mscomp_gpu_kernel_launches.argtypes = u64, [vp]
        L.mscomp_gpu_profile_enable.restype, L.mscomp_gpu_profile_enable.argtypes = None, [vp, i32]
        L.mscomp_gpu_set_encoder.restype, L.mscomp_gpu_set_encoder.argtypes = None, [vp, i32]
        L.mscomp_gpu_get_encoder.restype, L.mscomp_gpu_get_encoder.argtypes = i32, [vp]
        L.mscomp_gpu_profile_report.restype, L.mscomp_gpu_profile_report.argtypes = sz, [vp, vp, sz]
        L.mscomp_gpu_lznt1_compress.restype, L.mscomp_gpu_lznt1_compress.argtypes = i32, [vp, vp, sz, vp, sz, vp, vp]
        L.mscomp_gpu_lznt1_decompress.restype, L.mscomp_gpu_lznt1_decompress.argtypes = i32, [vp, vp, sz, vp, u32, vp, vp, vp, vp]
        L.mscomp_gpu_lznt1_index.restype, L.mscomp_gpu_lznt1_index.argtypes = i32, [vp, vp, sz, vp, sz, vp]
        for fmt in ("xpress_huff", "xpress"):
            for op in ("compress", "decompress"):
                if not hasattr(L, "mscomp_gpu_%s_%s_batch" % (fmt, op)):
                    continue
                f = getattr(L, "mscomp_gpu_%s_%s_batch" % (fmt, op))
                f.restype, f.argtypes = i32, [vp, vp, vp, vp]
        L.mscomp_gpu_xpress_huff_index_bound.restype, L.mscomp_gpu_xpress_huff_index_bound.argtypes = u64, [u64]
        L.mscomp_gpu_xpress_huff_index_batch.restype, L.mscomp_gpu_xpress_huff_index_batch.argtypes = i32, [vp, vp, vp, vp, vp, vp]
        L.mscomp_gpu_xpress_huff_decompress_batch_indexed.restype = i32
        L.mscomp_gpu_xpress_huff_decompress_batch_indexed.argtypes = [vp, vp, vp, vp, vp, vp, vp]
        L.mscomp_gpu_shard_range.restype, L.mscomp_gpu_shard_range.argtypes = None, [u64, u32, u32, vp, vp]
        L.mscomp_gpu_comm_unique_id.restype, L.mscomp_gpu_comm_unique_id.argtypes = i32, [vp]
        L.mscomp_gpu_comm_init.restype, L.mscomp_gpu_comm_init.argtypes = i32, [vp, vp, u32, u32]
        L.mscomp_gpu_comm_release.restype, L.mscomp_gpu_comm_release.argtypes = None, [vp]
        L.mscomp_gpu_comm_size.restype, L.mscomp_gpu_comm_size.argtypes = u32, [vp]
        L.mscomp_gpu_comm_rank.restype, L.mscomp_gpu_comm_rank.argtypes = u32, [vp]
        L.mscomp_gpu_sizes_allgather.restype, L.mscomp_gpu_sizes_allgather.argtypes = i32, [vp, vp, u32, vp]
        L.mscomp_gpu_collectives.restype, L.mscomp_gpu_collectives.argtypes = u64, [vp]
        L.mscomp_gpu_lznt1_compress_sharded.restype = i32
        L.mscomp_gpu_lznt1_compress_sharded.argtypes = [vp, vp, ctypes.c_size_t, vp, ctypes.c_size_t, vp, vp, vp]
        L.mscomp_gpu_compress_batch_sharded.restype, L.mscomp_gpu_compress_batch_sharded.argtypes = i32, [vp, i32, vp, vp, vp, vp, vp]
        L.mscomp_gpu_xpress_huff_compress_batch_indexed.restype = i32
        L.mscomp_gpu_xpress_huff_compress_batch_indexed.argtypes = [vp, vp, vp, vp, vp, vp, vp]
        L.mscomp_gpu_xpress_huff_second_chances.restype, L.mscomp_gpu_xpress_huff_second_chances.argtypes = u64, [vp]
        L.mscomp_gpu_xpress_compress_batch_indexed.restype = i32
        L.mscomp_gpu_xpress_compress_batch_indexed.argtypes = [vp, vp, vp, vp, u32, vp, u32]
        L.mscomp_gpu_xpress_decompress_batch_indexed.restype = i32
        L.mscomp_gpu_xpress_decompress_batch_indexed.argtypes = [vp, vp, vp, vp, vp, u32]
        _lib = L
    return _lib


# ---- host-pointer API, mirroring the reference's ctypes harness -------------------------------------------------

def _as_cbuf(data):
    """(address, length, keepalive) of a bytes-like object or a numpy / torch CPU uint8 array, without copying."""
    if hasattr(data, "data_ptr"):                       # torch tensor (host, possibly pinned)
        return data.data_ptr(), data.numel() * data.element_size(), data
    if hasattr(data, "ctypes") and hasattr(data, "nbytes"):   # numpy
        return data.ctypes.data, data.nbytes, data
    if isinstance(data, (bytes, bytearray)):
        buf = (ctypes.c_char * len(data)).from_buffer_copy(data) if isinstance(data, bytes) else (ctypes.c_char * len(data)).from_buffer(data)
        return ctypes.addressof(buf), len(data), buf
    mv = memoryview(data)
    buf = (ctypes.c_char * mv.nbytes).from_buffer_copy(mv.tobytes())
    return ctypes.addressof(buf), mv.nbytes, buf


class MSCompStream(ctypes.Structure):
    """mscomp_stream (include/mscomp/general.h; reference test/compressors.py:216-221)."""
    _fields_ = [("format", ctypes.c_int), ("compressing", ctypes.c_bool),
                ("in_", ctypes.c_void_p), ("in_avail", ctypes.c_size_t), ("in_total", ctypes.c_size_t),
                ("out", ctypes.c_void_p), ("out_avail", ctypes.c_size_t), ("out_total", ctypes.c_size_t),
                ("error", ctypes.c_char * 256), ("warning", ctypes.c_char * 256), ("state", ctypes.c_void_p)]


class OpenSrc:
    """Codec object of the same shape as the reference harness' `OpenSrc` (test/compressors.py:214-314): one-shot
    `Compress` / `Decompress` / `MaxCompressedSize` and the streaming members `CompressStream` / `DecompressStream`
    (a StreamableCompressor upstream, :224-229, :258-309).  `library` is a loaded libMSCompression (ctypes); the default is this
    package's GPU library -- the acceptance harness (tests/acceptance/) passes the compiled reference to get a second engine."""

    NO_FLUSH, FLUSH, FINISH = 0, 2, 4

    def __init__(self, format, library=None):
        self.format = int(format)
        self._library = library

    def _L(self):
        return self._library if self._library is not None else lib()

    def _call(self, name, *args):
        f = getattr(self._L(), name)
        f.restype = ctypes.c_int
        st = f(*args)
        if st < 0:
            raise MSCompError(name, st)
        return st

    def MaxCompressedSize(self, input_len):
        f = self._L().ms_max_compressed_size
        f.restype = ctypes.c_size_t
        val = f(ctypes.c_int(self.format), ctypes.c_size_t(input_len))
        if val == 0xFFFFFFFFFFFFFFFF:
            raise ValueError("invalid format")
        return val

    def Compress(self, input, output_buf=None):
        addr, n, keep = _as_cbuf(input)
        cap = self.MaxCompressedSize(n) + 16 if output_buf is None else len(output_buf)
        out = ctypes.create_string_buffer(max(cap, 1)) if output_buf is None else (ctypes.c_char * cap).from_buffer(output_buf)
        olen = ctypes.c_size_t(cap)
        self._call("ms_compress", ctypes.c_int(self.format), ctypes.c_void_p(addr), ctypes.c_size_t(n), ctypes.c_void_p(ctypes.addressof(out)), ctypes.byref(olen))
        return bytes(out[:olen.value])

    def Decompress(self, input, output_buf=None, out_size=None):
        addr, n, keep = _as_cbuf(input)
        if output_buf is None:
            cap = n * 4 if out_size is None else out_size
            out = ctypes.create_string_buffer(max(cap, 1))
        else:
            cap = len(output_buf)
            out = (ctypes.c_char * cap).from_buffer(output_buf)
        olen = ctypes.c_size_t(cap)
        self._call("ms_decompress", ctypes.c_int(self.format), ctypes.c_void_p(addr), ctypes.c_size_t(n), ctypes.c_void_p(ctypes.addressof(out)), ctypes.byref(olen))
        return bytes(out[:olen.value])

    # ---- streaming (reference test/compressors.py:258-309: the same loops, Python 3) ----
    def _pump(self, init, step, end, step_args, input, output, input_buf, output_buf):
        input_buf = bytearray(input_buf) if isinstance(input_buf, int) else (input_buf or bytearray(10 * 1024 * 1024))
        output_buf = bytearray(output_buf) if isinstance(output_buf, int) else (output_buf or bytearray(10 * 1024 * 1024))
        in_c = (ctypes.c_char * len(input_buf)).from_buffer(input_buf)
        out_c = (ctypes.c_char * len(output_buf)).from_buffer(output_buf)
        out_ptr, out_len = ctypes.addressof(out_c), len(output_buf)
        s = MSCompStream()
        self._call(init, ctypes.c_int(self.format), ctypes.byref(s))
        try:
            done = False
            s.in_avail = input.readinto(input_buf)
            while s.in_avail != 0:
                s.in_ = ctypes.addressof(in_c)
                while s.in_avail != 0:
                    s.out, s.out_avail = out_ptr, out_len
                    before = s.in_avail
                    done = self._call(step, ctypes.byref(s), *step_args(False)) >= 1
                    output.write(output_buf[:out_len - s.out_avail])
                    if s.in_avail == before and s.out_avail == out_len:
                        raise MSCompError(step, MSCOMP_BUF_ERROR)          # no progress with room on both sides: not a stream of this format
                s.in_avail = input.readinto(input_buf)
            for _ in range(1 << 20):
                if done:
                    break
                s.out, s.out_avail = out_ptr, out_len
                done = self._call(step, ctypes.byref(s), *step_args(True)) >= 1
                output.write(output_buf[:out_len - s.out_avail])
                if s.out_avail == out_len and not done:
                    break                                                 # nothing more comes out (a decoder that cannot tell the end)
        finally:
            self._call(end, ctypes.byref(s))

    def CompressStream(self, input, output, input_buf=None, output_buf=None):
        self._pump("ms_deflate_init", "ms_deflate", "ms_deflate_end",
                   lambda finishing: (ctypes.c_int(self.FINISH if finishing else self.NO_FLUSH),), input, output, input_buf, output_buf)

    def DecompressStream(self, input, output, input_buf=None, output_buf=None):
        self._pump("ms_inflate_init", "ms_inflate", "ms_inflate_end", lambda finishing: (), input, output, input_buf, output_buf)

    @property
    def streamable(self):
        """(compress, decompress): which streaming directions the library implements for this format (upstream: both for
        NoCompression and LZNT1, decompress only for Xpress, none for Xpress-Huffman; src/mscomp.cpp:134-223)."""
        return {CompressionFormat.NoCompression: (True, True), CompressionFormat.LZNT1: (True, True),
                CompressionFormat.Xpress: (False, True)}.get(self.format, (False, False))


OpenSrc.NoCompression = OpenSrc(CompressionFormat.NoCompression)
OpenSrc.LZNT1 = OpenSrc(CompressionFormat.LZNT1)
OpenSrc.Xpress = OpenSrc(CompressionFormat.Xpress)
OpenSrc.XpressHuffman = OpenSrc(CompressionFormat.XpressHuffman)


def raw_call(func_name, data, cap):
    """Call `<func_name>(in, in_len, out, &out_len)` of the library with an explicit capacity; returns (status, bytes|None)."""
    addr, n, keep = _as_cbuf(data)
    out = ctypes.create_string_buffer(max(cap, 1))
    olen = ctypes.c_size_t(cap)
    st = getattr(lib(), func_name)(addr, n, ctypes.addressof(out), ctypes.byref(olen))
    return st, (bytes(out[:olen.value]) if st == 0 else None)


# ---- device-resident API over torch tensors ------------------------------------------------------------------------

RESULT_BYTES = 24


def parse_result(t):
    """mscomp_gpu_result (include/mscomp_gpu.h) from a 24-byte uint8 tensor; synchronises through .cpu()."""
    b = bytes(t.cpu().numpy().tobytes())
    nbytes, status, first_bad, short_units, _ = struct.unpack("<QiIII", b[:RESULT_BYTES])
    return {"bytes": nbytes, "status": status, "first_bad": first_bad, "short_units": short_units}


class Batch(ctypes.Structure):
    """mscomp_gpu_batch (include/mscomp_gpu.h)."""
    _fields_ = [("count", ctypes.c_uint32), ("d_in_off", ctypes.c_void_p), ("d_in_len", ctypes.c_void_p),
                ("d_out_off", ctypes.c_void_p), ("d_out_cap", ctypes.c_void_p), ("d_out_len", ctypes.c_void_p),
                ("d_status", ctypes.c_void_p)]


class GpuContext:
    """mscomp_gpu_ctx bound to a torch device and to torch's CURRENT stream (so torch.cuda.Event timing sees the kernels)."""

    def __init__(self, device=None):
        import torch
        self.torch = torch
        if not torch.cuda.is_available():
            raise RuntimeError("mscompress_b200.GpuContext needs a CUDA device (there is no CPU fallback)")
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        with torch.cuda.device(self.device):
            self.stream = torch.cuda.current_stream()
            self.handle = lib().mscomp_gpu_create_on_stream(self.device.index, ctypes.c_void_p(self.stream.cuda_stream))
        if not self.handle:
            raise RuntimeError("mscomp_gpu_create failed")

    def close(self):
        if getattr(self, "handle", None):
            lib().mscomp_gpu_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def launches(self):
        return int(lib().mscomp_gpu_kernel_launches(self.handle))

    # ---- several GPUs (include/mscomp_gpu.h "several GPUs"): the library's own NCCL communicator ----
    def comm_init(self, rank, world, broadcast):
        """Collective.  Rank 0 makes the 128-byte NCCL id; `broadcast(bytes_or_None) -> bytes` hands it to every rank (any
        transport: torch.distributed, MPI, a file); then ncclCommInitRank inside the library."""
        idbuf = ctypes.create_string_buffer(128)
        if rank == 0:
            st = lib().mscomp_gpu_comm_unique_id(ctypes.addressof(idbuf))
            if st < 0:
                raise MSCompError("mscomp_gpu_comm_unique_id", st)
        got = broadcast(bytes(idbuf.raw) if rank == 0 else None)
        idbuf = ctypes.create_string_buffer(bytes(got), 128)
        st = lib().mscomp_gpu_comm_init(self.handle, ctypes.addressof(idbuf), rank, world)
        if st < 0:
            raise MSCompError("mscomp_gpu_comm_init", st)

    def comm_size(self):
        return int(lib().mscomp_gpu_comm_size(self.handle))

    def collectives(self):
        return int(lib().mscomp_gpu_collectives(self.handle))

    def sizes_allgather(self, d_mine, d_all):
        """d_all[r * n + k] = rank r's d_mine[k] (int64 / uint64 device tensors), one ncclAllGather on this context's stream."""
        st = lib().mscomp_gpu_sizes_allgather(self.handle, d_mine.data_ptr(), d_mine.numel(), d_all.data_ptr())
        if st < 0:
            raise MSCompError("mscomp_gpu_sizes_allgather", st)

    def lznt1_compress_sharded(self, d_in, d_out, d_chunk_off, d_result, d_sizes_all):
        """This rank's shard of an LZNT1 stream through the single-GPU path + the all-gather of the per-rank compressed sizes."""
        st = lib().mscomp_gpu_lznt1_compress_sharded(self.handle, d_in.data_ptr(), d_in.numel(), d_out.data_ptr(), d_out.numel(),
                                                     d_chunk_off.data_ptr(), d_result.data_ptr(), d_sizes_all.data_ptr())
        if st < 0:
            raise MSCompError("mscomp_gpu_lznt1_compress_sharded", st)

    def profile(self, on=True):
        lib().mscomp_gpu_profile_enable(self.handle, 1 if on else 0)

    def set_encoder(self, mode):
        """"reference" (default: byte-identical to the reference's Xpress / Xpress-Huffman encoders) or "fast"."""
        lib().mscomp_gpu_set_encoder(self.handle, {"reference": 0, "fast": 1}[mode])

    def encoder(self):
        return ("reference", "fast")[lib().mscomp_gpu_get_encoder(self.handle)]

    def profile_report(self):
        """{kernel_name: (calls, total_ms)} since profiling was enabled / last reported (synchronises)."""
        buf = ctypes.create_string_buffer(1 << 16)
        lib().mscomp_gpu_profile_report(self.handle, ctypes.addressof(buf), len(buf))
        out = {}
        for line in buf.value.decode().splitlines():
            name, calls, ms = line.split()
            out[name] = (int(calls), float(ms))
        return out

    def synchronize(self):
        st = lib().mscomp_gpu_synchronize(self.handle)
        if st < 0:
            raise MSCompError("mscomp_gpu_synchronize", st)

    def _u8(self, n):
        return self.torch.empty(n, dtype=self.torch.uint8, device=self.device)

    # -- LZNT1 stream --
    def lznt1_compress(self, d_in, d_out=None, d_chunk_off=None, d_result=None):
        t = self.torch
        n = d_in.numel()
        nchunks = (n + 4095) // 4096
        if d_out is None:
            d_out = self._u8(n + 2 * nchunks + 16)
        if d_chunk_off is None:
            d_chunk_off = t.empty(nchunks + 1, dtype=t.int64, device=self.device)
        if d_result is None:
            d_result = self._u8(RESULT_BYTES)
        st = lib().mscomp_gpu_lznt1_compress(self.handle, d_in.data_ptr(), n, d_out.data_ptr(), d_out.numel(),
                                             d_chunk_off.data_ptr(), d_result.data_ptr())
        if st < 0:
            raise MSCompError("mscomp_gpu_lznt1_compress", st)
        return d_out, d_chunk_off, d_result

    def lznt1_decompress(self, d_in, d_chunk_off, nchunks, d_out=None, d_result=None, d_sizes=None, d_status=None, in_len=None):
        if d_out is None:
            d_out = self._u8(nchunks * 4096)
        if d_result is None:
            d_result = self._u8(RESULT_BYTES)
        st = lib().mscomp_gpu_lznt1_decompress(self.handle, d_in.data_ptr(), d_in.numel() if in_len is None else in_len,
                                               d_chunk_off.data_ptr(), nchunks, d_out.data_ptr(),
                                               d_sizes.data_ptr() if d_sizes is not None else None,
                                               d_status.data_ptr() if d_status is not None else None, d_result.data_ptr())
        if st < 0:
            raise MSCompError("mscomp_gpu_lznt1_decompress", st)
        return d_out, d_result

    def lznt1_index(self, d_in, chunk_cap=None, in_len=None):
        """Chunk index of an unindexed LZNT1 stream, built on the GPU.  Returns (offsets tensor, info dict with nchunks,
        consumed, tail); tail: 0 clean, 1 stray zero byte, 2 stray byte, 3 terminator, 4 terminator + data, 5 truncated,
        6 bad signature."""
        n = d_in.numel() if in_len is None else in_len
        if chunk_cap is None:
            chunk_cap = n // 3 + 1
        d_off = self.torch.empty(chunk_cap + 1, dtype=self.torch.int64, device=self.device)
        res = (ctypes.c_uint64 * 3)()
        st = lib().mscomp_gpu_lznt1_index(self.handle, d_in.data_ptr(), n, d_off.data_ptr(), chunk_cap, ctypes.addressof(res))
        info = {"nchunks": int(res[0]), "consumed": int(res[1]), "tail": int(res[2] & 0xFFFFFFFF)}
        if st < 0 and st != -5:
            raise MSCompError("mscomp_gpu_lznt1_index", st)
        info["status"] = st
        return d_off, info

    # -- batches of independent buffers (Xpress-Huffman, Xpress) --
    def make_batch(self, in_off, in_len, out_off, out_cap):
        """Device descriptor arrays from host sequences; returns (Batch, out_len tensor, status tensor, keepalive)."""
        t = self.torch
        dev = self.device
        a = t.as_tensor(in_off, dtype=t.int64).to(dev)
        b = t.as_tensor(in_len, dtype=t.int32).to(dev)
        c = t.as_tensor(out_off, dtype=t.int64).to(dev)
        d = t.as_tensor(out_cap, dtype=t.int32).to(dev)
        olen = t.zeros(len(in_len), dtype=t.int32, device=dev)
        stat = t.zeros(len(in_len), dtype=t.int32, device=dev)
        batch = Batch(len(in_len), a.data_ptr(), b.data_ptr(), c.data_ptr(), d.data_ptr(), olen.data_ptr(), stat.data_ptr())
        return batch, olen, stat, (a, b, c, d)

    def run_batch(self, fmt, op, d_in, d_out, batch):
        """fmt in {'xpress_huff','xpress'}, op in {'compress','decompress'}."""
        f = getattr(lib(), "mscomp_gpu_%s_%s_batch" % (fmt, op))
        st = f(self.handle, d_in.data_ptr(), d_out.data_ptr(), ctypes.byref(batch))
        if st < 0:
            raise MSCompError(f.__name__, st)

    # -- Xpress-Huffman restart index (mscomp_gpu_xhuff_seg: 4 x uint32 per entry) --
    def xhuff_index_alloc(self, out_caps):
        """(index tensor, offsets tensor [count+1], nseg tensor) with room for streams of the given output capacities."""
        t = self.torch
        bound = [int(lib().mscomp_gpu_xpress_huff_index_bound(int(c))) for c in out_caps] if len(out_caps) < 4096 else None
        if bound is None:                           # large batches: the bound formula, vectorised (cap/64 + cap/65536 + 3)
            caps = t.as_tensor(out_caps, dtype=t.int64)
            b = caps // 64 + caps // 65536 + 3
        else:
            b = t.as_tensor(bound, dtype=t.int64)
        off = t.zeros(len(out_caps) + 1, dtype=t.int64)
        off[1:] = t.cumsum(b, 0)
        total = int(off[-1])
        return (t.zeros(max(total, 1) * 4, dtype=t.int32, device=self.device), off.to(self.device),
                t.zeros(len(out_caps), dtype=t.int32, device=self.device))

    def xhuff_index(self, d_in, batch, d_index, d_index_off, d_nseg):
        """The serial walk only: fills the batch's out_len / status and the restart index (no output bytes)."""
        st = lib().mscomp_gpu_xpress_huff_index_batch(self.handle, d_in.data_ptr(), ctypes.byref(batch), d_index.data_ptr(),
                                                      d_index_off.data_ptr(), d_nseg.data_ptr())
        if st < 0:
            raise MSCompError("mscomp_gpu_xpress_huff_index_batch", st)

    def xhuff_decompress_indexed(self, d_in, d_out, batch, d_index, d_index_off, d_nseg):
        """Decode from a restart index kept by the caller: one warp per stream, one lane per 64-symbol segment."""
        st = lib().mscomp_gpu_xpress_huff_decompress_batch_indexed(self.handle, d_in.data_ptr(), d_out.data_ptr(), ctypes.byref(batch),
                                                                   d_index.data_ptr(), d_index_off.data_ptr(), d_nseg.data_ptr())
        if st < 0:
            raise MSCompError("mscomp_gpu_xpress_huff_decompress_batch_indexed", st)

    def xhuff_compress_indexed(self, d_in, d_out, batch, d_index, d_index_off, d_nseg):
        """Xpress-Huffman encode that also leaves the side-band restart index of every stream (the streams equal those of run_batch)."""
        st = lib().mscomp_gpu_xpress_huff_compress_batch_indexed(self.handle, d_in.data_ptr(), d_out.data_ptr(), ctypes.byref(batch),
                                                                 d_index.data_ptr(), d_index_off.data_ptr(), d_nseg.data_ptr())
        if st < 0:
            raise MSCompError("mscomp_gpu_xpress_huff_compress_batch_indexed", st)

    def xhuff_second_chances(self):
        """Streams the indexed Xpress-Huffman decoder handed to the index-free decoder so far (diagnostic)."""
        return int(lib().mscomp_gpu_xpress_huff_second_chances(self.handle))

    XPRESS_SEG_WORDS = 6     # uint32 fields of mscomp_gpu_xpress_seg

    def xpress_index_tensor(self, count, stride):
        """Device storage for the restart index of `count` plain-Xpress buffers, `stride` entries each."""
        return self.torch.zeros(count * stride * self.XPRESS_SEG_WORDS, dtype=self.torch.int32, device=self.device)

    def xpress_compress_indexed(self, d_in, d_out, batch, seg_bytes, d_index, stride):
        """Plain Xpress encode that also writes the side-band restart index (the streams equal those of run_batch)."""
        st = lib().mscomp_gpu_xpress_compress_batch_indexed(self.handle, d_in.data_ptr(), d_out.data_ptr(), ctypes.byref(batch), seg_bytes,
                                                            d_index.data_ptr(), stride)
        if st < 0:
            raise MSCompError("mscomp_gpu_xpress_compress_batch_indexed", st)

    def xpress_decompress_indexed(self, d_in, d_out, batch, d_index, stride):
        """Decode of streams made by xpress_compress_indexed: one lane per segment instead of one lane per buffer."""
        st = lib().mscomp_gpu_xpress_decompress_batch_indexed(self.handle, d_in.data_ptr(), d_out.data_ptr(), ctypes.byref(batch),
                                                              d_index.data_ptr(), stride)
        if st < 0:
            raise MSCompError("mscomp_gpu_xpress_decompress_batch_indexed", st)


# ---- synthetic data ---------------------------------------------------------------------------------

def _genlib():
    global _gen
    if _gen is None:
        if not os.path.exists(GEN_PATH):
            raise ImportError("mscompress_b200: %s is missing -- run the build" % GEN_PATH)
        g = ctypes.CDLL(GEN_PATH)
        for name in ("msgen_text", "msgen_random", "msgen_runs"):
            getattr(g, name).restype = None
            getattr(g, name).argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint64]
        g.msgen_batch.restype = None
        g.msgen_batch.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_uint64, ctypes.c_int]
        _gen = g
    return _gen


KINDS = {"text": 0, "random": 1, "runs": 2}


def generate(kind, n, seed=42):
    """n bytes of generator `kind` ('text' | 'random' | 'runs') as a bytes object."""
    buf = ctypes.create_string_buffer(max(n, 1))
    getattr(_genlib(), "msgen_" + kind)(ctypes.addressof(buf), n, seed)
    return buf.raw[:n]


def generate_batch_into(array, each, count, seed0=42, kind="text", threads=None):
    """Fill a writable contiguous uint8 numpy array with `count` buffers of `each` bytes (buffer i: seed0+i); threaded."""
    import concurrent.futures as cf
    g = _genlib()
    base = array.ctypes.data
    threads = threads or min(32, os.cpu_count() or 1)
    per = (count + threads - 1) // threads
    jobs = [(i, min(per, count - i)) for i in range(0, count, per)]

    def run(job):
        i, c = job
        g.msgen_batch(base + i * each, each, c, seed0 + i, KINDS[kind])
    with cf.ThreadPoolExecutor(threads) as ex:
        list(ex.map(run, jobs))
    return array
