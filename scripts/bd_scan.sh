timeout 300 python -m pytest tests/test_gpu_block.py tests/test_gpu_zz_block_iface.py -x -q 2>&1 | tail -1
scripts/blk_scan.sh C5-2M bd0 bd1p | tail -3
B200_BLK_NO_PACK=1 scripts/blk_scan.sh C5-2M bd0 | tail -1
scripts/blk_scan.sh C5 bd0 bd1p | tail -3
