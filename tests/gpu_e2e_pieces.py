import sys, time, ctypes as C
import numpy as np
sys.path.insert(0,'.')
import __graft_entry__ as g
import hashlib, json
corpus=g.load_submodule('corpus'); pkg=g.load_package(); L=pkg.lib()
data=np.frombuffer(corpus.generate_cached('C5'),dtype=np.uint8)
n,W=len(data),8192
px,pl=L.x3s_host_alloc(n+W),L.x3s_host_alloc(n)
x=np.ctypeslib.as_array(C.cast(px,C.POINTER(C.c_uint8)),shape=(n+W,)); x[:n]=data; x[n:]=0
out=np.ctypeslib.as_array(C.cast(pl,C.POINTER(C.c_uint8)),shape=(n,))
tm=pkg.Timing()
for mb in (32,16,8):
    import os; os.environ['X3_SEG_PIECE_MB']=str(mb)
    best=1e9
    for r in range(6):
        t0=time.perf_counter(); assert L.x3s_search_host(px,n,W,15,1,0,pl,None,C.byref(tm))==0; best=min(best,(time.perf_counter()-t0)*1e3)
    sha=hashlib.sha256(out.tobytes()).hexdigest()
    want=json.load(open('tests/golden/tables.json'))['C5']['lstar_sha256']
    print(f'piece {mb} MB: host-to-host {best:.2f} ms, kernel part {tm.kernel_ms:.2f}, launches {tm.launches}, table ok {sha==want}',flush=True)
