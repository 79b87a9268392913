set +e
export D377_TIMELINE=1
run() { echo "== $*"; env "$@" timeout 200 python tools/tune_msm.py $L 2>&1 | grep "^n=\|group"; }
{
L=24
run D377_X=0
run D377_MSM_GW=3,8,3
run D377_MSM_GW=2,5,5,2
run D377_MSM_GW=3,9,2
run D377_MSM_GW=4,8,2
run D377_MSM_GW=2,10,2
L=20
run D377_X=0
run D377_MSM_GW=4,8,4
run D377_MSM_GW=5,8,3
run D377_MSM_GW=8,8
run D377_MSM_GW=10,6
run D377_MSM_GW=6,7,3
L=22
run D377_X=0
run D377_MSM_GW=3,8,3
run D377_MSM_GW=4,8,2
run D377_MSM_GW=8,6
} > gpurun_out/s4g_tune.log 2>&1
sed -E 's/run=auto seg=auto: //; s/scan=0.00. scatter=0.00. //; s/stitch=0.00. bucket_reduce=0.00. //' gpurun_out/s4g_tune.log
