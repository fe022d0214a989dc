source scripts/sweep.sh
run leaf_nc HM_X=1
run leaf_cs HM_LIB=$V/libhairmsnn_leafcs.so
run leaf_evict_first HM_LIB=$V/libhairmsnn_leafef.so
run leaf_nc_again HM_X=1
