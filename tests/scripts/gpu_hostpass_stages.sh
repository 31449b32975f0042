python - <<'PY'
import sys
sys.path.insert(0,'.')
import __graft_entry__ as g
corpus=g.load_submodule('corpus')
open('/tmp/c5_32m.bin','wb').write(corpus.generate('C5', 32_000_000))
PY
nproc; lscpu | grep -E "Model name|Thread|Core|Socket" | head -5
for th in -4 4 2 1; do echo "X3_THREADS=$th"; X3_THREADS=$th x3-compressor_b200/bin/x3 -zf /tmp/c5_32m.bin /tmp/c5_32m.x3 2>&1 | grep -E "elapsed|stages|start-up"; done
