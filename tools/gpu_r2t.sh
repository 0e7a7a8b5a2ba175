#!/bin/bash
# round 2, call T: importance sampling (RT_FLAG_IMPORTANCE) parity + convergence; compute-sanitizer attempt
set -x
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 900 python -m pytest tests/test_importance.py -m gpu -q > $O/r2t_pytest.log 2>&1; tail -15 $O/r2t_pytest.log
timeout 300 python tools/ab_probe.py --variants 4 --cases "7:1024x1024x64,8:1024x1024x64,9:1920x1080x32" --flags 0,0x800 --tag importance > $O/r2t_ab.jsonl 2> $O/r2t_ab.err
cat $O/r2t_ab.jsonl | cut -c1-250
timeout 240 compute-sanitizer --tool racecheck python tools/sanitize_run.py > $O/r2t_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -5 $O/r2t_racecheck.log
