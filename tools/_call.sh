timeout 900 python -m pytest tests/test_gpu_eas.py -x -q -m gpu -k "displacement_gradient" 2>&1 | tail -4
python tools/config_times.py C4dg 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print(d['config'][:60], d['elements_ms'], d['K_R_Melem_s'])"
