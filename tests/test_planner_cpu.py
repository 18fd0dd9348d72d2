"""Host logic of the planner on CPU: the reference's constructor arguments (planner/planner.py:214-227), per-rank capacities
(dynamicemb_config.py:696-765), the plan entries the sharders consume."""
import pytest


class _Cfg:
    def __init__(self, name, dim, num):
        self.name, self.embedding_dim, self.num_embeddings, self.feature_names = name, dim, num, [f"f_{name}"]


class _Topology:
    world_size = 8


def _cons():
    from dynamicemb import DynamicEmbScoreStrategy, DynamicEmbTableOptions
    from dynamicemb.shard import DynamicEmbParameterConstraints
    mk = lambda **kw: DynamicEmbTableOptions(score_strategy=DynamicEmbScoreStrategy.STEP, **kw)      # noqa: E731
    return {"user": DynamicEmbParameterConstraints(use_dynamicemb=True, dynamicemb_options=mk(dist_type="hash_roundrobin")),
            "item": DynamicEmbParameterConstraints(use_dynamicemb=True, dynamicemb_options=mk(bucket_capacity=64)),
            "dense_tbl": DynamicEmbParameterConstraints(use_dynamicemb=False)}


def test_planner_reference_constructor_and_capacities():
    from dynamicemb import get_sharded_table_capacity
    from dynamicemb.shard import DynamicEmbeddingShardingPlanner
    cfgs = [_Cfg("user", 128, 1_000_003), _Cfg("item", 128, 10_000_000_000), _Cfg("dense_tbl", 64, 1000)]
    cons = _cons()
    planner = DynamicEmbeddingShardingPlanner(eb_configs=cfgs, topology=_Topology(), batch_size=512, constraints=cons, debug=False)
    plan = planner.collective_plan(module=None, sharders=[], pg=None)
    assert set(plan) == {"user", "item"}                                      # tables without use_dynamicemb are TorchRec's business
    assert plan["user"]["local_capacity"] == get_sharded_table_capacity(1_000_003, 8, 128) == 125_056      # round_up(ceil(N / 8), 128)
    assert plan["item"]["local_capacity"] == get_sharded_table_capacity(10_000_000_000, 8, 64) == 1_250_000_000
    for name, dist_type in (("user", "hash_roundrobin"), ("item", "roundrobin")):
        e = plan[name]
        assert e["sharding_type"] == "row_wise" and e["compute_kernel"] == "customized_kernel" and e["ranks"] == list(range(8))
        assert e["dist_type"] == dist_type and e["dynamicemb_options"] is cons[name].dynamicemb_options
        assert cons[name].dynamicemb_options.max_capacity == e["local_capacity"]                          # written back, as the reference does
    assert planner.plan(cfgs, []) == plan                                    # reference call form plan(module, sharders): positional non-dict ignored


def test_planner_earlier_call_form_and_errors():
    from dynamicemb.shard import DynamicEmbeddingShardingPlanner
    cons = _cons()
    plan = DynamicEmbeddingShardingPlanner(cons, world_size=2).plan({"user": 50_000, "item": 200_000})
    assert plan["user"]["local_capacity"] == 25_088 and plan["item"]["local_capacity"] == 100_032 and plan["item"]["ranks"] == [0, 1]
    with pytest.raises(ValueError, match="no num_embeddings"):
        DynamicEmbeddingShardingPlanner(constraints=_cons(), world_size=2).plan()
    assert DynamicEmbeddingShardingPlanner(constraints=_cons()).plan({"user": 128, "item": 64})["user"]["ranks"] == [0]      # no process group: one rank
