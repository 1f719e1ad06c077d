set -x
python profiles/prof_trunk.py 4096 3 2>&1 | tail -1
for mb in 30 45 60 75 105; do MZB_STACK_STICKY_MB=$mb python profiles/prof_trunk.py 4096 3 2>&1 | tail -1; done
MZB_STACK_STICKY_MB=45 MZB_STACK_OTHER_POLICY=1 python profiles/prof_trunk.py 4096 3 2>&1 | tail -1
MZB_STACK_STICKY_MB=105 MZB_STACK_STICKY_FRAC=0.5 python profiles/prof_trunk.py 4096 3 2>&1 | tail -1
python profiles/prof_trunk.py 4096 3 2>&1 | tail -1
