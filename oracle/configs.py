"""The five BASELINE.json configs, resolved as SURVEY.md section 8 "config resolution" prescribes.

Each entry: reference yml + the minimal documented overrides that make the reference classes
compose (never by editing reference sources), + the OracleCfg fields they imply.
TEST INFRASTRUCTURE (oracle/), also read by bench.py's cpu_baseline / --impl reference legs.
"""
from .ekl_oracle import OracleCfg

# name -> (reference yml, cfg overrides applied after the yml, oracle fields)
CONFIGS = {
    # 1: trainer.py semantics, COND_G_NET(E+1+1024) + JOINT_D_NET64/128; CAT_Z 'sum' affects dims only
    "catcls": dict(yml="birds_2stgs_catcls.yml", batch=24,
                   over={"TRAIN.CAT_Z": "sum"},
                   ocfg=dict(G_KIND="cond", COND="txt+cls", CLS_KIND="multihot", CAT_Z="sum", Z_DIM=100)),
    # 2: 3-branch COND_G_NET + JOINT_D_NET64/128/256 (both reference trainers assert on BRANCH_NUM>2)
    "3stages": dict(yml="birds_3stages.yml", batch=24,
                    over={"TRAIN.CAT_Z": "sum"},
                    ocfg=dict(G_KIND="cond", COND="txt+cls", CLS_KIND="multihot", CAT_Z="sum", Z_DIM=100,
                              BRANCH_NUM=3)),
    # 3: capsule-conditioned G/D: COND_G_NET(1024, use_cap) (cub:135) + JOINT_D(use_cap)
    "onlycapsule": dict(yml="birds_2stgs_onlycapsule.yml", batch=32,
                        over={"TRAIN.CAT_Z": "sum", "TRAIN.G_CAPSULE": True, "TRAIN.D_CAPSULE": True},
                        ocfg=dict(G_KIND="cond", COND="txt", CLS_KIND="index", CAT_Z="sum", Z_DIM=100,
                                  G_CAPSULE=True, D_CAPSULE=True)),
    # 4: runs as written through cub_trainer_splitz_cap_ca.py
    "splitz_cap_ca": dict(yml="birds_2stg_splitz_cap_ca.realcls.yml", batch=32, over={},
                          ocfg=dict(G_KIND="catz_ca", CLS_KIND="index", CAT_Z="concat", Z_DIM=128,
                                    G_CAPSULE=True, D_CAPSULE=True)),
    # 5: as 1 with ENTITY_DIM 90
    "coco": dict(yml="coco_2stgs.yml", batch=64,
                 over={"TRAIN.CAT_Z": "sum"},
                 ocfg=dict(G_KIND="cond", COND="txt+cls", CLS_KIND="multihot", CAT_Z="sum", Z_DIM=100,
                           ENTITY_DIM=90)),
    # ---- conditioning variants of config 4 (one cfg override each), pinned by the real reference too
    # CAT_Z sum / product (model.py:500-505, cub:577-582)
    "splitz_cat_sum": dict(yml="birds_2stg_splitz_cap_ca.realcls.yml", batch=32, over={"TRAIN.CAT_Z": "sum"},
                           ocfg=dict(G_KIND="catz_ca", CLS_KIND="index", CAT_Z="sum", Z_DIM=128, G_CAPSULE=True, D_CAPSULE=True)),
    "splitz_cat_product": dict(yml="birds_2stg_splitz_cap_ca.realcls.yml", batch=32, over={"TRAIN.CAT_Z": "product"},
                               ocfg=dict(G_KIND="catz_ca", CLS_KIND="index", CAT_Z="product", Z_DIM=128, G_CAPSULE=True,
                                         D_CAPSULE=True)),
    # TREE.SCALE 4: upsample2 in NEXT_STAGE_G (model.py:406-407, 420-421), JOINT_D_NET256 as the stage-2 discriminator
    # (cub:151-154); composable only with a non-concat CAT_Z because JOINT_D_NET256 ignores CAT_Z (model.py:1210)
    "splitz_scale4_sum": dict(yml="birds_2stg_splitz_cap_ca.realcls.yml", batch=32, over={"TRAIN.CAT_Z": "sum", "TREE.SCALE": 4},
                              ocfg=dict(G_KIND="catz_ca", CLS_KIND="index", CAT_Z="sum", SCALE=4, Z_DIM=128, G_CAPSULE=True,
                                        D_CAPSULE=True)),
    # COND_G_NET_CATZ (model.py:567-665: two VC_NETs; no shipped trainer builds it): with the exchange capsule stem
    # (COND_INIT_STAGE_G_Exchange_Cap, model.py:280-333) and with the plain Linear stem, driven through the cub step
    "catz_exchange": dict(yml="birds_2stg_splitz_cap_ca.realcls.yml", batch=32, over={"TRAIN.EXCHANGE": True},
                          ocfg=dict(G_KIND="catz", CLS_KIND="index", CAT_Z="concat", EXCHANGE=True, Z_DIM=128, G_CAPSULE=True,
                                    D_CAPSULE=True)),
    "catz_plain": dict(yml="birds_2stg_splitz_cap_ca.realcls.yml", batch=32, over={"TRAIN.G_CAPSULE": False, "TRAIN.D_CAPSULE": False},
                       ocfg=dict(G_KIND="catz", CLS_KIND="index", CAT_Z="concat", Z_DIM=128, G_CAPSULE=False, D_CAPSULE=False)),
    # config 2's generator + the two-head D_NET64/128/256.  NOT pinned by a golden fixture: the reference's step functions
    # cannot drive these modules (app. A #14); the modules themselves are pinned (tests/golden/modules.npz dnet*).
    "3stages_dnet": dict(yml="birds_3stages.yml", batch=24, over={"TRAIN.CAT_Z": "sum"},
                         ocfg=dict(G_KIND="cond", COND="txt+cls", CLS_KIND="multihot", CAT_Z="sum", Z_DIM=100, BRANCH_NUM=3)),
}
NO_REFERENCE_STEP = ("3stages_dnet",)      # configs the reference cannot run: no golden step fixture

BASELINE = ("catcls", "3stages", "onlycapsule", "splitz_cap_ca", "coco")      # the five BASELINE.json configs


def oracle_cfg(name, batch=None, gf=None, df=None):
    spec = CONFIGS[name]
    kw = dict(spec["ocfg"])
    kw["BATCH_SIZE"] = batch or spec["batch"]
    if gf:
        kw["GF_DIM"] = gf
    if df:
        kw["DF_DIM"] = df
    return OracleCfg(**kw)


def cond_dim(c):
    """cond width of COND_G_NET's VC_NET: trainer.py:116 (E+1+text) or cub:135 (text)."""
    return c.TEXT_DIM + c.ENTITY_DIM + 1 if c.COND == "txt+cls" else c.TEXT_DIM
