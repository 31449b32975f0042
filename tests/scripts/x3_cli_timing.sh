python - <<'PY'
import sys
sys.path.insert(0,'.')
import __graft_entry__ as g
c = g.load_submodule("corpus")
d = c.generate("C2")
open('/tmp/c2','wb').write(d); open('/tmp/c2_1m','wb').write(d[:1000000]); open('/tmp/c2_3m','wb').write(d[:3000000])
PY
for f in c2_1m c2_3m c2 c2_1m; do echo == $f; ./x3-compressor_b200/bin/x3 -zf /tmp/$f /tmp/$f.x3 2>&1 | grep -E "elapsed|search|events|entries"; done
