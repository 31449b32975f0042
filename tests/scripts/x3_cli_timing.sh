python - <<'PY'
import sys
sys.path.insert(0,'.')
import __graft_entry__ as g
c = g.load_submodule("corpus")
d = c.generate("C2")
open('/tmp/c2','wb').write(d); open('/tmp/c2_1m','wb').write(d[:1000000])
PY
for f in c2_1m c2 c2; do echo == $f; X3_TRACE=1 ./x3-compressor_b200/bin/x3 -zf /tmp/$f /tmp/$f.x3 2>&1 | grep -E "elapsed|search|x3s_search_host|x3k_launch_rank: (scratch|[0-9])"; done
for f in c2; do echo == $f stream kernel; X3_SEARCH_KERNEL=3 X3_TRACE=1 ./x3-compressor_b200/bin/x3 -zf /tmp/$f /tmp/$f.x3 2>&1 | grep -E "elapsed|search|x3s_search_host"; done
