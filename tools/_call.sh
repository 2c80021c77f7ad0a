export QT_NO_PCG=1
python tools/quick_time.py 2>&1 | grep -E "^elements|^step"
IKB_H8_BULK=0 python tools/quick_time.py 2>&1 | grep -E "^elements|^step"
