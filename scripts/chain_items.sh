export PYTHONUNBUFFERED=1
for it in 5 200 450 800 1000 1200 1270; do DXI_LIB=$PWD/deepxi_b200/libdeepxi_b200_dbg.so timeout -k 5 120 python scripts/chain_clocks.py 256 10 $it 2>&1 | grep -E "block period|GEMM1 done|aux landed|P2.1 done" | tr '\n' ' '; echo; done
