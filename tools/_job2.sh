for l in 2200 2600 3000; do run exp_libs/lib_lag$l.so lag$l h2o20 0; done
