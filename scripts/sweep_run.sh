source scripts/sweep.sh
run split8_span6 HM_BVH_SPLIT=8 HM_BVH_SPAN=6
run split16_span5 HM_BVH_SPLIT=16 HM_BVH_SPAN=5
run split16_span3 HM_BVH_SPLIT=16 HM_BVH_SPAN=3
run split12_span8 HM_BVH_SPLIT=12 HM_BVH_SPAN=8
run s8s10_p12s8 HM_BVH_SPLIT=8 HM_BVH_SPAN=10 HM_LIB=$V/libhairmsnn_p12s8.so
run s8s10_p8s6 HM_BVH_SPLIT=8 HM_BVH_SPAN=10 HM_LIB=$V/libhairmsnn_p8s6.so
run s8s10_p16s10n6 HM_BVH_SPLIT=8 HM_BVH_SPAN=10 HM_LIB=$V/libhairmsnn_p16s10n6.so
run s8s10_p16s10 HM_BVH_SPLIT=8 HM_BVH_SPAN=10 HM_LIB=$V/libhairmsnn_p16s10.so
