run() {
    echo "== $*"
    env "$@" SWS_B200_DEBUG=1 python tools/bench_configs.py --only "C4,X2,X3,X1,E2" 2>&1 | grep -E "scale8:|^C4|^X|^E" | sort -u | cut -c1-215
}
run A=1
run SWS_B200_S8_TH=16
run SWS_B200_S8_TH=8
run SWS_B200_S8_KB=56
run SWS_B200_S8_KB=56 SWS_B200_S8_TH=16
run SWS_B200_S8_KB=44
run SWS_B200_S8_KB=100
run SWS_B200_S8_KB=56 SWS_B200_S8_STAGES=3
