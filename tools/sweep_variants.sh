python tools/trace_solve.py cfg2 2>/dev/null | tail -1
LSA_NO_SUBTREES=1 python tools/trace_solve.py cfg2 2>/dev/null | tail -1
python tools/trace_solve.py cfg1 2>/dev/null | tail -1
LSA_NO_SUBTREES=1 python tools/trace_solve.py cfg1 2>/dev/null | tail -1
python tools/trace_solve.py cav3d 2>/dev/null | tail -1
LSA_NO_SUBTREES=1 python tools/trace_solve.py cav3d 2>/dev/null | tail -1
