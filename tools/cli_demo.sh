python - <<'PY'
import sys
sys.path.insert(0, ".")
from aletsch_b200 import hostlib as H
cfg = H.default_config(H.SYNTH_PAIRED, chrom_len=20_000_000, seed=7)
syn = H.Synth(cfg)
for k in range(3):
    H.write_bam("/tmp/demo%d.bam" % k, syn.sample(k, 300000, threads=8), [cfg.chrom_len] * cfg.n_chrom)
PY
ls -la /tmp/demo*.bam
python -m aletsch_b200.run --clusters --max-group-size 20 --min-grouping-similarity 0.2 /tmp/demo0.bam /tmp/demo1.bam /tmp/demo2.bam
