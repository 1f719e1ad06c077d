python profiles/prof_trunk.py 4096 3 2>&1 | tail -1
MZB_STACK_DEBUG=8 python profiles/prof_trunk.py 4096 3 2>&1 | tail -1
MZB_STACK_DEBUG=16 python profiles/prof_trunk.py 4096 3 2>&1 | tail -1
python profiles/prof_trunk.py 4096 3 2>&1 | tail -1
MZB_STACK_DEBUG=8 python profiles/prof_trunk.py 4096 3 2>&1 | tail -1
python profiles/prof_trunk.py 1024 2 2>&1 | tail -1
MZB_STACK_DEBUG=8 python profiles/prof_trunk.py 1024 2 2>&1 | tail -1
python profiles/prof_trunk.py 256 2 2>&1 | tail -1
MZB_STACK_DEBUG=8 python profiles/prof_trunk.py 256 2 2>&1 | tail -1
