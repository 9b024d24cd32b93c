#!/bin/bash
# round-end trip: what the driver runs (validate.sh), then the ncu launch list and --set full captures for profiles/
tools/gpu/validate.sh
COUNTK=2 tools/gpu/profiles.sh > gpurun_out/profiles_trip.log 2>&1; tail -n 8 gpurun_out/profiles_trip.log
