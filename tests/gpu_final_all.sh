#!/bin/bash
# End of round: the checked build against the whole GPU suite, then tests/gpu_final_check.sh (smoke, suite, both arms).
bash tests/gpu_checked_build.sh
bash tests/gpu_final_check.sh
