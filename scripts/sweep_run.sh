source scripts/sweep.sh
run tail37 HM_X=1
run tail8 HM_LIB=$V/libhairmsnn_tail8.so
run tail16 HM_LIB=$V/libhairmsnn_tail16.so
run tail24 HM_LIB=$V/libhairmsnn_tail24.so
run pairs HM_TAIL_MEGA=0
