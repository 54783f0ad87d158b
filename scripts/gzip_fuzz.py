"""Randomised files through the host's gzip readers against zlib (no GPU needed): mixed data (FASTQ-like, random, runs,
short periods), compression levels and strategies, memory levels, Z_SYNC_FLUSH / Z_FULL_FLUSH blocks in the middle,
one to three members per file; read through `merkurio records <file> cat` with the sequential reader or the parallel
one (2-8 threads, pieces of 4-500 KiB) in read sizes of 1 000 bytes to 8 MiB. The soak test beside
tests/test_inflate_cpu.py; it found the one bug of the decoder (profiles/sanitizer/README.md).

    python scripts/gzip_fuzz.py [seed] [seconds]
"""
import gzip
import os
import random
import subprocess
import sys
import tempfile
import time
import zlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, 'merkurio_b200', 'lib', 'merkurio')
TMP = tempfile.mkdtemp(prefix='mk_gzfuzz_')
rng = random.Random(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
def fastq(n):
    out=[]
    for i in range(n):
        L=rng.randint(30,300)
        s=''.join(rng.choices('ACGTN',weights=[30,20,20,30,1],k=L)); q=''.join(rng.choices('FFFF:,#IJ5',k=L))
        out.append(f'@r{i} {rng.random()}\n{s}\n+\n{q}\n')
    return ''.join(out).encode()
def mix():
    parts=[]
    for _ in range(rng.randint(1,6)):
        k=rng.random()
        if k<0.4: parts.append(fastq(rng.randint(500,8000)))
        elif k<0.55: parts.append(os.urandom(rng.randint(1000,400000)))
        elif k<0.7: parts.append(bytes([rng.randrange(256)])*rng.randint(1000,3000000))
        elif k<0.85: parts.append((os.urandom(rng.randint(1,40)))*rng.randint(100,50000))
        else: parts.append(b''.join(bytes([rng.randrange(4)+65])*rng.randint(1,12) for _ in range(rng.randint(1000,100000))))
    return b''.join(parts)
def comp(raw):
    lvl=rng.choice([1,1,3,6,6,9,0]); strat=rng.choice([zlib.Z_DEFAULT_STRATEGY]*4+[zlib.Z_FILTERED,zlib.Z_HUFFMAN_ONLY,zlib.Z_RLE,zlib.Z_FIXED]); mem=rng.choice([9,8,4,1])
    c=zlib.compressobj(lvl,zlib.DEFLATED,31,mem,strat)
    out=b''
    pos=0
    while pos<len(raw):
        n=rng.randint(1,max(1,len(raw)//3)); out+=c.compress(raw[pos:pos+n]); pos+=n
        if rng.random()<0.2: out+=c.flush(rng.choice([zlib.Z_SYNC_FLUSH,zlib.Z_FULL_FLUSH]))  # empty stored blocks in the middle
    return out+c.flush()
bad=0; t0=time.time(); it=0
while time.time() - t0 < (float(sys.argv[2]) if len(sys.argv) > 2 else 300):
    it+=1
    members=[mix() for _ in range(rng.choice([1,1,1,2,3]))]
    raw=b''.join(members); data=b''.join(comp(m) for m in members)
    open(os.path.join(TMP, 'fz.gz'),'wb').write(data)
    env={**os.environ,'MERKURIO_GZIP_THREADS':str(rng.choice([1,1,2,3,5,8])),'MERKURIO_GZIP_PIECE_KB':str(rng.choice([4,5,8,16,33,64,128,500]))}
    r=subprocess.run([BIN,'records',os.path.join(TMP, 'fz.gz'),'cat',str(rng.choice([1000,70000,8<<20]))],capture_output=True,env=env)
    if r.returncode!=0 or r.stdout!=raw:
        bad+=1; print('MISMATCH',it,len(raw),len(data),env['MERKURIO_GZIP_THREADS'],env['MERKURIO_GZIP_PIECE_KB'],r.returncode,r.stderr[:200]); open(os.path.join(TMP, f'fail{it}.gz'), 'wb').write(data)
print('iterations', it, 'bad', bad, '(failing files, if any, are kept in ' + TMP + ')')
if bad == 0:
    import shutil
    shutil.rmtree(TMP, ignore_errors=True)
