source scripts/sweep.sh
run s16_5 HM_X=1
run s16_4 HM_BVH_SPAN=4
run s16_7 HM_BVH_SPAN=7
run s8_8 HM_BVH_SPLIT=8 HM_BVH_SPAN=8
