# Selector block of the reference's examples/active/cbgs_spatial_temporal.py:361-367, stand-alone.
# Paths follow the reference's layout; override with your own.
data_root = "data/nuScenes"
train_anno = data_root + "/infos_train_10sweeps_withvelo.pkl"

selector = dict(
    type="SpatialTemporalSelector",
    budget=4800,                                   # placeholder: --budget always overrides it
    buffer_file=data_root + "/buffer_spatial_temporal.json",
    infos_origin=train_anno,
    logs_file=data_root + "/v1.0-trainval/log.json",
    # reference constructor defaults, spelled out (README.md:39-40 tells users to edit the source):
    k=8, normalize="exp", aggregate="sum", lambda_t=1, cost_b=0.04, cost_f=0.12,
)
