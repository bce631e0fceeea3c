#!/bin/bash
# How the DATA fixtures under tests/golden/ were taken from the reference checkout (run in the build container, where
# /root/reference exists; the GPU box has no reference - the tests read only these copies).  No reference SOURCE is copied.
set -e
REF=${1:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
mkdir -p "$HERE/data" "$HERE/learned" "$HERE/logs" "$HERE/models/params_0.2_learnable" "$HERE/models/params_1.0_fixed_one-one"
# initial gating matrices (utils_data.py:147-176) with their pandas CSV exports, and the full-data co-occurrence table
cp "$REF"/data/gating_matrix_*.npy "$REF"/data/gating_matrix_*.csv "$REF"/data/label_cooccurance_matrix.csv "$HERE/data/"
# the same five .npy files also live in <repo>/data/ (the reference's own location, utils_data.py:149): bench.py, smoke() and scripts/ read them there
mkdir -p "$HERE/../../data" && cp "$REF"/data/gating_matrix_*.npy "$HERE/../../data/"
# learned gating matrices (gated_ccvae.py:396-403)
for f in 0.2 0.5 1.0; do for m in best last; do
  cp "$REF/models/params_${f}_learnable/learned_gating_matrix_${m}.npy" "$HERE/learned/learned_gating_matrix_${f}_${m}.npy"
done; done
# one complete trained checkpoint (Keras-2.8 save_weights files + learned mu) and the small files of a one-one model
cp "$REF"/models/params_0.2_learnable/{encoder_model_best.h5,decoder_model_best.h5,classifier_best.h5,cond_prior_best.h5,learned_gating_matrix_best.npy,learned_gating_matrix_best.csv} "$HERE/models/params_0.2_learnable/"
cp "$REF"/models/params_1.0_fixed_one-one/{cond_prior_best.h5,classifier_best.h5} "$HERE/models/params_1.0_fixed_one-one/"
# the 75 logged temperature values of a learnable run (gated_ccvae.py:404-406)
grep -h "decayed" "$REF/models/params_1.0_learnable/logs" | sed 's/.*decayed to: //' > "$HERE/logs/gating_sampler_temp_params_1.0_learnable.txt"
chmod -R u+w "$HERE"
