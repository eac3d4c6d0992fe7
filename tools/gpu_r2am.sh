#!/bin/bash
timeout 200 python tools/eager_period.py
GCA_PREP_LATE=1 timeout 200 python tools/eager_period.py
