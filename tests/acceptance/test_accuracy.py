#!/usr/bin/env python
"""
test_accuracy.py format directory

Python 3 port of the reference's acceptance harness (test/test_accuracy.py:38-95): every file of `directory` goes through every
(compressor engine, decompressor engine) pair of the chosen format -- one of none, lznt1, xpress, xpress-huffman -- one-shot and,
where the format streams (test/compressors.py:224-229), through the streaming members with 100 KiB + 1 byte buffers (:52, :87).
No news is good news: only errors and a few status lines are printed; the exit status is the number of errors.

As a pytest module (`-m gpu`) the same matrix runs over a corpus made by the three synthetic generators of SURVEY.md 8(d);
MSCOMP_ACCEPTANCE_MIB sets the size of the large file of each kind (default 4 MiB; 1024 = the 1 GiB of BASELINE.json).
"""
import io
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import compressors as C                                          # noqa: E402

STREAM_BUF = 100 * 1024 + 1


def check_file(fullpath, data, engines, errors, log=None):
    def err(msg):
        errors.append(msg)
        print("Error: " + msg, file=sys.stderr)

    def decompress(compressed, name1, codec, name2, how):
        try:
            if how == "one-shot":
                decomp = codec.Decompress(compressed, out_size=len(data))
            else:
                sink = io.BytesIO()
                codec.DecompressStream(io.BytesIO(compressed), sink, STREAM_BUF, STREAM_BUF)
                decomp = sink.getvalue()
            if len(data) != len(decomp):
                err("%s failed to %s %s-decompress %s compressed data (length %d != %d)" % (fullpath, name2, how, name1, len(data), len(decomp)))
            elif data != decomp:
                err("%s failed to %s %s-decompress %s compressed data (content mismatch)" % (fullpath, name2, how, name1))
        except Exception as ex:
            err("%s failed to %s %s-decompress %s compressed data (%s)" % (fullpath, name2, how, name1, ex))

    def fan_out(compressed, name1):
        for name2, codec2 in engines.items():
            decompress(compressed, name1, codec2, name2, "one-shot")
            if codec2.streamable[1]:
                decompress(compressed, name1, codec2, name2, "stream")

    for name1, codec1 in engines.items():
        try:
            fan_out(codec1.Compress(data), name1)
        except Exception as ex:
            err("%s failed to %s compress (%s)" % (fullpath, name1, ex))
        if codec1.streamable[0]:
            try:
                sink = io.BytesIO()
                codec1.CompressStream(io.BytesIO(data), sink, STREAM_BUF, STREAM_BUF)
                fan_out(sink.getvalue(), name1 + " (stream)")
            except Exception as ex:
                err("%s failed to %s stream-compress (%s)" % (fullpath, name1, ex))


def run(fmt, path, need_gpu=True):
    engines = C.engines_for(fmt, need_gpu)
    if engines is None or not os.path.isdir(path):
        print("Error: arguments", file=sys.stderr)
        return 2
    print("The available compressors are: " + ", ".join(engines))
    errors = []
    t0 = time.perf_counter()
    for root, dirs, files in os.walk(path):
        print("%8.2f Folder: %s" % (time.perf_counter() - t0, root))
        for name in sorted(files):
            fullpath = os.path.join(root, name)
            try:
                with open(fullpath, "rb") as f:
                    data = f.read()
            except OSError:
                continue
            if len(data) == 0:
                continue
            check_file(fullpath, data, engines, errors)
    print("%8.2f Done" % (time.perf_counter() - t0))
    return len(errors)


# ---- pytest entry: the generators' corpus ------------------------------------------------------------------------
try:
    import pytest
except ImportError:                                              # the script form does not need it
    pytest = None

if pytest is not None:
    @pytest.fixture(scope="module")
    def corpus(tmp_path_factory):
        import mscompress_b200 as ms
        d = tmp_path_factory.mktemp("corpus")
        big = int(os.environ.get("MSCOMP_ACCEPTANCE_MIB", "4")) << 20
        for kind in ("text", "random", "runs"):
            for n in (1, 4095, 4096, 65536, 65537, 300001, big if kind == "text" else min(big, 1 << 20)):
                with open(os.path.join(d, "%s_%d.bin" % (kind, n)), "wb") as f:
                    f.write(ms.generate(kind, n, 42 + n % 13))
        return str(d)

    @pytest.mark.gpu
    @pytest.mark.parametrize("fmt", ["none", "lznt1", "xpress", "xpress-huffman"])
    def test_every_engine_pair_round_trips_the_corpus(fmt, corpus):
        assert run(fmt, corpus) == 0

    @pytest.mark.parametrize("fmt", ["none", "lznt1", "xpress", "xpress-huffman"])
    def test_harness_on_the_reference_engine_alone(fmt, tmp_path):
        """No GPU: the harness itself (loops, streaming members, buffer sizes) against the compiled reference."""
        import mscompress_b200 as ms
        if not C.load_libraries(need_gpu=False):
            pytest.skip("compiled reference not available")
        for kind, n in (("text", 300001), ("random", 4097), ("runs", 70000), ("text", 1)):
            (tmp_path / ("%s_%d.bin" % (kind, n))).write_bytes(ms.generate(kind, n, 42))
        assert run(fmt, str(tmp_path), need_gpu=False) == 0


if __name__ == "__main__":
    if len(sys.argv) != 3:
        print("Error: arguments", file=sys.stderr)
        sys.exit(2)
    sys.exit(min(run(sys.argv[1], sys.argv[2]), 125))
