python scripts/_capture_stress.py plain 2>&1 | grep -v "^lpm-b200" | tail -8
python scripts/_capture_stress.py nogc 2>&1 | grep -v "^lpm-b200" | tail -8
LPM_FUSE_OPT=0 python scripts/_capture_stress.py plain 2>&1 | grep -v "^lpm-b200" | tail -4
