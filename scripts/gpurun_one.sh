# usage: gpurun -- 'bash scripts/gpurun_one.sh <pytest args>'   (one focused test selection)
mkdir -p gpurun_out
timeout 900 python -m pytest "$@" -q -x -m gpu > gpurun_out/pytest_one.log 2>&1; echo "pytest exit $?"; tail -30 gpurun_out/pytest_one.log
