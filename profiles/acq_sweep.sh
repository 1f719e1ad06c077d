for n in 256 1024 4096; do
python profiles/prof_trunk.py $n 2 2>&1 | tail -1
MZB_STACK_DEBUG=32 python profiles/prof_trunk.py $n 2 2>&1 | tail -1
done
