for cfg in "0 0 0:1" "100 288 0:4" "80 408 0:4" "110 228 0:4" "120 168 0:4" "100 288 0:8" "148 0 0:4"; do
  set -- $cfg
  echo "### scatter_ctas=$1 probe_ctas=$2 config=$3"
  sc=$1; pc=$2
  if [ "$sc" != "0" ]; then export PHJ_DIST_SCATTER_CTAS=$sc; else unset PHJ_DIST_SCATTER_CTAS; fi
  if [ "$pc" != "0" ]; then export PHJ_DIST_PROBE_CTAS=$pc; else unset PHJ_DIST_PROBE_CTAS; fi
  CONFIGS=$3 JOINS=5 timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/probe_dist.py 2>&1 | grep -v "^\*\*\*\|OMP_NUM\|^$\|NCCL version"
done
