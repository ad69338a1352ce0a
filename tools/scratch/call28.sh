cd $GRAFT_REPO_ROOT
R="timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
echo whole BASE; $R --master-port 29521 tools/ddp_check.py BASE_CFG 2>&1 | grep -E "rel-l2" | head -1
echo perseg BASE; VCD_BWD_WHOLE=0 $R --master-port 29522 tools/ddp_check.py BASE_CFG 2>&1 | grep -E "rel-l2" | head -1
echo whole SMALL; $R --master-port 29523 tools/ddp_check.py 2>&1 | grep -E "rel-l2" | head -1
echo perseg SMALL; VCD_BWD_WHOLE=0 $R --master-port 29524 tools/ddp_check.py 2>&1 | grep -E "rel-l2" | head -1
echo whole BASE nographs; VCD_GRAPHS=0 $R --master-port 29525 tools/ddp_check.py BASE_CFG 2>&1 | grep -E "rel-l2" | head -1
