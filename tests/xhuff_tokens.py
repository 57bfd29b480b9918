"""TEST INFRASTRUCTURE: a plain-Python reader of Xpress-Huffman streams that returns the LZ77 TOKENS of every chunk (MS-XCA 2.2;
the reference's decoder is src/xpress_huff_decompress.cpp:39-135).  Used to compare an encoder's parse with the reference's
independently of the Huffman code lengths it chose.  Small inputs only."""


def tokens(comp, out_len):
    """-> [chunk][token]; token = int (literal byte) or (length, offset)."""
    chunks, ip, produced = [], 0, 0
    while produced < out_len or not chunks:
        lens = []
        for b in comp[ip:ip + 256]:
            lens += [b & 15, b >> 4]
        ip += 256
        # canonical decode table
        codes, code = {}, 0
        for n in range(1, 16):
            for s in range(512):
                if lens[s] == n:
                    codes[(n, code)] = s
                    code += 1
            code <<= 1
        # bit reader (Bitstream.h:34-107): 16-bit little-endian words, MSB first
        def word(at):
            return comp[at] | (comp[at + 1] << 8) if at + 1 < len(comp) else 0
        mask = (word(ip) << 16) | word(ip + 2)
        bits, ip = 32, ip + 4
        toks, chunk_end = [], min(produced + 65536, out_len)
        while True:
            if produced >= chunk_end and not (chunk_end == out_len and False):
                break
            n, c, sym = 0, 0, None
            while sym is None:
                n += 1
                c = (c << 1) | ((mask >> (32 - n)) & 1)
                sym = codes.get((n, c))
                assert n <= 15
            mask = (mask << n) & 0xFFFFFFFF
            bits -= n
            if bits < 16:
                mask |= word(ip) << (16 - bits)
                bits += 16
                ip += 2
            if sym < 256:
                toks.append(sym)
                produced += 1
                continue
            sym -= 256
            if sym == 0 and produced >= out_len:
                break
            ln, ob = sym & 15, sym >> 4
            if ln == 15:
                ln = comp[ip]
                ip += 1
                if ln == 255:
                    ln = comp[ip] | (comp[ip + 1] << 8)
                    ip += 2
                    assert ln != 0
                    ln -= 15
                ln += 15
            ln += 3
            off = ((mask >> (32 - ob)) if ob else 0) + (1 << ob)
            mask = (mask << ob) & 0xFFFFFFFF
            bits -= ob
            if bits < 16:
                mask |= word(ip) << (16 - bits)
                bits += 16
                ip += 2
            toks.append((ln, off))
            produced += ln
        chunks.append(toks)
        if produced >= out_len:
            break
        # the chunk's bit stream ends where the reader stands: the words it prefetched belong to it (the encoder reserves them)
    return chunks
