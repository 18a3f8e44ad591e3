for cfg in "128 0" "128 1" "128 2" "64 0" "64 1" "64 2" "64 3"; do set -- $cfg; echo "== tile $1 variant $2" >> gpurun_out/r02t_ab.txt; NARDE_TILE=$1 NARDE_VARIANT=$2 python tools/timeline_probe.py 131072 0 2>&1 | tail -2 >> gpurun_out/r02t_ab.txt; done
cat gpurun_out/r02t_ab.txt
NARDE_TILE=64 python tools/phase_clock.py 131072 2>&1 | head -12 > gpurun_out/r02t_phase64.txt; cat gpurun_out/r02t_phase64.txt
NARDE_TILE=128 python tools/phase_clock.py 131072 2>&1 | head -12 > gpurun_out/r02t_phase128.txt; cat gpurun_out/r02t_phase128.txt
python -m pytest tests/test_gpu_actor.py tests/test_game_manager.py -m gpu -x -q > gpurun_out/r02t_pytest.log 2>&1; tail -5 gpurun_out/r02t_pytest.log
