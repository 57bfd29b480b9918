"""Driven by tests/test_emu_sanitized.py in a subprocess with libasan/libubsan preloaded: every kernel, every format,
on tests/emu/libemu_san.so (the kernels compiled with -fsanitize=address,undefined)."""
import os
import sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, ROOT)
import mscompress_b200 as ms, emu_util
emu=ctypes.CDLL(os.path.join(ROOT, 'tests', 'emu', 'libemu_san.so'))
orc=ctypes.CDLL(os.path.join(ROOT, 'oracle', 'liboracle.so'))
def ocall(name,data,cap):
    out=ctypes.create_string_buffer(max(cap,1)); n=ctypes.c_size_t(cap)
    rc=getattr(orc,'orc_'+name)(data,ctypes.c_size_t(len(data)),out,ctypes.byref(n)); return rc,out.raw[:n.value]
MAX={"xpress": lambda n: n + 4 + 4 * (n // 32), "xpress_huff": lambda n: n + 34 + 258 + 258 * (n // 65536)}
plains=[ms.generate(k,n,5) for k in ('text','runs','random') for n in (1,100,5000,70000,140000)]
for fmt,(dk,ek) in {"xpress_huff":("emu_xhuff_decode","emu_xhuff_encode"),"xpress":("emu_xpress_decode","emu_xpress_encode")}.items():
    st,ol,outs=emu_util.emu_batch(emu,ek,plains,[MAX[fmt](len(p)) for p in plains])
    assert not st.any()
    for p,c in zip(plains,outs): assert ocall(fmt+'_decompress',c,len(p))==(0,p)
    st,ol,outs2=emu_util.emu_batch(emu,dk,outs,[len(p) for p in plains])
    assert not st.any() and outs2==plains
    print(fmt,'ok')
for kern in ("emu_lznt1_decode",):
    for p in plains:
        rc,c=ocall('lznt1_compress',p,len(p)+100)
        st,sz,ch=emu_util.emu_lznt1_decode(emu,c,kernel=kern); assert not st.any() and b''.join(ch)==p
        e,_=emu_util.emu_lznt1_encode(emu,p); assert e==c
    print(kern,'ok')
# the reference-parse encoders (lz_exact.cuh): byte-identical to the oracle's compressors; independent buffers and the one-shot stream layout
xp=plains+[bytes(9000)+plains[3], bytes(30000)+b"x"+bytes(30000)+plains[2], plains[3]*3]
st,_,outs=emu_util.emu_batch(emu,"emu_xpress_encode_exact",xp,[MAX["xpress"](len(p)) for p in xp],grid=3,extra=(ctypes.c_uint32(16384),ctypes.c_void_p(0),ctypes.c_uint32(0)))
assert not st.any()
for p,c in zip(xp,outs): assert c==ocall('xpress_compress',p,MAX["xpress"](len(p))+16)[1]
st,_,outs=emu_util.emu_batch(emu,"emu_xhuff_encode_exact",xp,[MAX["xpress_huff"](len(p)) for p in xp],grid=3,extra=(ctypes.c_uint32(0),))
assert not st.any()
for p,c in zip(xp,outs): assert c==ocall('xpress_huff_compress',p,MAX["xpress_huff"](len(p))+16)[1]
for p in xp:
    if len(p)>65536:
        st,c=emu_util.emu_xhuff_stream_exact(emu,p); assert not st.any() and c==ocall('xpress_huff_compress',p,MAX["xpress_huff"](len(p))+16)[1]
print('reference-parse encoders ok')
# two-kernel decoders, warp parser + block apply, restart index, heavy LZNT1 chunks, GPU chunk index
import numpy as np, random
for fmt,dk in (("xpress_huff","emu_xhuff_decode2"),("xpress","emu_xpress_decode2"),("xpress","emu_xpress_decode_warp")):
    comps=[ocall(fmt+'_compress',p,MAX[fmt](len(p))+16)[1] for p in plains]
    st,ol,outs2=emu_util.emu_batch(emu,dk,comps,[len(p) for p in plains])
    assert not st.any() and outs2==plains
    print(dk,'ok')
stride=12; index=np.zeros(len(plains)*stride*6,dtype=np.uint32)
st,_,comps=emu_util.emu_batch(emu,"emu_xpress_encode_indexed",plains,[MAX["xpress"](len(p)) for p in plains],extra=(ctypes.c_uint32(16384),emu_util.P(index),ctypes.c_uint32(stride)))
assert not st.any()
for blk in (0,1):
    st,_,outs2=emu_util.emu_batch(emu,"emu_xpress_decode_indexed",comps,[len(p) for p in plains],extra=(emu_util.P(index),ctypes.c_uint32(stride),ctypes.c_uint32(blk)))
    assert not st.any() and outs2==plains
print('xpress restart index ok')
rnd=random.Random(3)
heavy=bytes(rnd.choice(b"ACGT") for _ in range(2*4096+50))
rc,c=ocall('lznt1_compress',heavy,len(heavy)+100)
for hw in (32768,0):
    e,_=emu_util.emu_lznt1_encode(emu,heavy,heavy_work=hw); assert e==c
rc,c=ocall('lznt1_compress',plains[3],len(plains[3])+100)
a,addr=emu_util.padded(c); offs=np.zeros(len(c)//3+3,dtype=np.uint64); stt=np.zeros(4,dtype=np.uint64)
emu.emu_lznt1_index.restype=ctypes.c_int
assert emu.emu_lznt1_index(ctypes.c_void_p(addr),ctypes.c_uint64(len(c)),ctypes.c_uint32(4112),ctypes.c_uint32(3),emu_util.P(offs),ctypes.c_uint64(len(c)//3+2),emu_util.P(stt))==0
assert int(stt[0])==1 and int(stt[3])==(len(plains[3])+4095)//4096
print('heavy chunks + chunk index ok')
# one plain Xpress stream by speculation (xpress_spec.cuh): groups at every offset, block tables, true chain, restart index, indexed parse, pointer jumping
emu.emu_xpress_stream_decode.restype=ctypes.c_uint32
for p in (plains[3],plains[8],plains[13],plains[4]*2):
    c=ocall('xpress_compress',p,MAX["xpress"](len(p))+16)[1]
    for comp,cap in ((c,len(p)),(c[:-3],len(p)),(c,len(p)-1)):
        a,addr=emu_util.padded(comp); out=np.zeros(cap+128,dtype=np.uint8); tot=ctypes.c_uint64(0); info=(ctypes.c_uint32*4)()
        h=emu.emu_xpress_stream_decode(ctypes.c_void_p(addr),ctypes.c_uint32(len(comp)),ctypes.c_void_p(out.ctypes.data+64),ctypes.c_uint64(cap),ctypes.byref(tot),info)
        if comp is c and cap==len(p): assert h==1 and out[64:64+tot.value].tobytes()==p
        else: assert h==0 or (0,out[64:64+tot.value].tobytes())==ocall('xpress_decompress',comp,cap)
print('xpress speculative stream decode ok')
# one Xpress-Huffman stream: candidates, a warp per chunk (xhuff_chunk_warp.cuh), pointer jumping -- multi-chunk and single-chunk, whole and truncated
emu.emu_xhuff_stream_decode.restype=ctypes.c_uint32
for p in (plains[4],plains[3],plains[9],plains[12]):
    c=ocall('xpress_huff_compress',p,MAX["xpress_huff"](len(p))+16)[1]
    for comp,cap in ((c,len(p)),(c[:-3],len(p)),(c,len(p)-1)):
        a,addr=emu_util.padded(comp); out=np.zeros(cap+128,dtype=np.uint8); tot=ctypes.c_uint64(0); info=ctypes.c_uint32(0)
        h=emu.emu_xhuff_stream_decode(ctypes.c_void_p(addr),ctypes.c_uint32(len(comp)),ctypes.c_void_p(out.ctypes.data+64),ctypes.c_uint64(cap),ctypes.byref(tot),ctypes.byref(info))
        if comp is c and cap==len(p): assert h==1 and out[64:64+tot.value].tobytes()==p
        else: assert h==0 or (0,out[64:64+tot.value].tobytes())==ocall('xpress_huff_decompress',comp,cap)
print('xpress-huffman chunk-parallel stream decode ok')
print('sanitized run ok')
