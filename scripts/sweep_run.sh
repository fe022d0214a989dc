source scripts/sweep.sh
run base HM_X=1
run aw1 HM_LIB=$V/libhairmsnn_aw1.so
run aw2 HM_LIB=$V/libhairmsnn_aw2.so
run aw1cw2 HM_LIB=$V/libhairmsnn_aw1cw2.so
run aw1cw1 HM_LIB=$V/libhairmsnn_aw1cw1.so
run aw1p12 HM_LIB=$V/libhairmsnn_aw1p12.so
