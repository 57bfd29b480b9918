"""Re-entrancy of the C ABI ("the shim must be callable concurrently"): every host thread gets its own context
(stream, scratch, pinned mailbox); concurrent calls of all six codec functions must return exactly what they return alone."""
import threading

import pytest
from conftest import MAXSZ

pytestmark = pytest.mark.gpu


def test_concurrent_calls_from_many_threads(ms, oracle):
    work = []
    for t in range(8):
        fmt = ("lznt1", "xpress", "xpress_huff")[t % 3]
        n = (5 << 20) + 4097 * t if fmt == "lznt1" else 200000 + 777 * t
        plain = ms.generate(("text", "runs", "random")[t % 3], n, 100 + t)
        work.append((fmt, plain, oracle.compress(fmt, plain)))
    errors = []

    def run(fmt, plain, ref_comp):
        try:
            for _ in range(3):
                st, comp = ms.raw_call(fmt + "_compress", plain, MAXSZ[fmt](len(plain)))
                assert st == 0
                if fmt == "lznt1":
                    assert comp == ref_comp                      # byte-identical encoder
                assert oracle.decompress(fmt, comp, len(plain)) == (0, plain)
                assert ms.raw_call(fmt + "_decompress", ref_comp, len(plain)) == (0, plain)
                assert ms.raw_call(fmt + "_decompress", ref_comp, len(plain) - 1)[0] < 0
        except Exception as e:  # noqa: BLE001
            errors.append((fmt, repr(e)))

    threads = [threading.Thread(target=run, args=w) for w in work]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors
