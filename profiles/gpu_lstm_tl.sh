#!/bin/bash
# rebuild the library WITH the clock64 timeline (debug build, not the shipped one), run the LSTM timeline, then rebuild without
HCA_BUILD_TIMELINE=1 python visual-question-answering_b200/build.py --force > /dev/null 2>&1; echo "debug build rc=$?"
timeout 200 python profiles/timeline_lstm.py 2>&1 | tail -22
