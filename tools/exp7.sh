python tools/msm_loop.py 25 26 2>&1 | grep -v stages
python tools/msm_loop.py --fast -c=18 26 2>&1 | grep -v stages
python tools/msm_loop.py --fast -c=20 26 2>&1 | grep -v stages
python tools/msm_loop.py --fast -c=21 25 2>&1 | grep -v stages
python tools/msm_loop.py --fast -c=19 25 2>&1 | grep -v stages
