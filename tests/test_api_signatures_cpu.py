"""The drop-in boundary, checked on CPU: every function this package offers under a reference name takes the reference's parameters — same
names, same order, same literal defaults — as recorded in tests/golden/api_signatures.json (read from the reference's source by
tests/golden/gen_golden_sigs.py).  A caller written against the reference, positional or by keyword, binds identically here."""
import inspect
import json
import os

import pytest

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "api_signatures.json")


def _ours():
    from dynamicemb import BatchedDynamicEmbeddingTablesV2 as M
    from dynamicemb import FrequencyAdmissionStrategy as FA, KVCounter, LinearBucketTable as LT, MultiTableKVCounter as MC, get_scored_table
    from dynamicemb import checkpoint as ck, dump_load as dl, types as ty
    from hstu import fused_hstu_op as fo
    from hstu import hstu_attn_varlen_func, hstu_ops_gpu as HO, layer_ops as L
    return {
        "fused_hstu_op": fo.fused_hstu_op, "FusedHSTULayerFunction.forward": fo.FusedHSTULayerFunction.forward,
        "triton_weighted_layer_norm_fwd": L.triton_weighted_layer_norm_fwd, "triton_weighted_layer_norm_bwd": L.triton_weighted_layer_norm_bwd,
        "triton_layer_norm_mul_dropout_fwd": L.triton_layer_norm_mul_dropout_fwd, "triton_layer_norm_mul_dropout_bwd": L.triton_layer_norm_mul_dropout_bwd,
        "triton_silu_fwd": L.triton_silu_fwd, "triton_silu_bwd": L.triton_silu_bwd, "hstu_attn_varlen_func": hstu_attn_varlen_func,
        "hstu_varlen_fwd_100": HO.hstu_varlen_fwd_100, "hstu_varlen_bwd_100": HO.hstu_varlen_bwd_100,
        "BatchedDynamicEmbeddingTablesV2.__init__": M.__init__, "BatchedDynamicEmbeddingTablesV2.forward": M.forward,
        "BatchedDynamicEmbeddingTablesV2.prefetch": M.prefetch, "BatchedDynamicEmbeddingTablesV2.dump": M.dump, "BatchedDynamicEmbeddingTablesV2.load": M.load,
        "BatchedDynamicEmbeddingTablesV2.export_keys_values": M.export_keys_values, "BatchedDynamicEmbeddingTablesV2.set_score": M.set_score,
        "BatchedDynamicEmbeddingTablesV2.set_learning_rate": M.set_learning_rate,
        "BatchedDynamicEmbeddingTablesV2.incremental_dump": M.incremental_dump, "BatchedDynamicEmbeddingTablesV2.fill_tables": M.fill_tables,
        "BatchedDynamicEmbeddingTablesV2.split_embedding_weights": M.split_embedding_weights,
        "BatchedDynamicEmbeddingTablesV2.reset_cache_states": M.reset_cache_states,
        "BatchedDynamicEmbeddingTablesV2.set_record_cache_metrics": M.set_record_cache_metrics,
        "BatchedDynamicEmbeddingTablesV2.flush": M.flush, "BatchedDynamicEmbeddingTablesV2.get_score": M.get_score,
        "KVCounter.__init__": KVCounter.__init__, "MultiTableKVCounter.__init__": MC.__init__, "MultiTableKVCounter.add": MC.add,
        "MultiTableKVCounter.erase": MC.erase, "MultiTableKVCounter.memory_usage": MC.memory_usage, "MultiTableKVCounter.load": MC.load,
        "MultiTableKVCounter.dump": MC.dump, "FrequencyAdmissionStrategy.__init__": FA.__init__, "FrequencyAdmissionStrategy.admit": FA.admit,
        "FrequencyAdmissionStrategy.initialize_non_admitted_embeddings": FA.initialize_non_admitted_embeddings,
        "LinearBucketTable.__init__": LT.__init__, "LinearBucketTable.lookup": LT.lookup, "LinearBucketTable.insert": LT.insert,
        "LinearBucketTable.insert_and_evict": LT.insert_and_evict, "LinearBucketTable.erase": LT.erase, "LinearBucketTable.load": LT.load,
        "LinearBucketTable.dump": LT.dump, "get_scored_table": get_scored_table,
        "get_planner": __import__("dynamicemb.get_planner", fromlist=["get_planner"]).get_planner,
        **{n: getattr(__import__("dynamicemb.benchmark.dataset_generator", fromlist=[n]), n) for n in ("translateToPowerLaw", "PowerLaw", "gen_key", "zipf")},
        "find_sharded_modules": dl.find_sharded_modules, "get_dynamic_emb_module": dl.get_dynamic_emb_module, "DynamicEmbDump": dl.DynamicEmbDump,
        "DynamicEmbLoad": dl.DynamicEmbLoad, "set_score": dl.set_score, "get_score": dl.get_score, "incremental_dump": dl.incremental_dump,
        "is_valid_score_threshold": dl.is_valid_score_threshold, "get_sharded_table_capacity": ty.get_sharded_table_capacity,
        "get_table_value_bytes": ty.get_table_value_bytes, "string_to_evict_strategy": ty.string_to_evict_strategy,
        "dyn_emb_to_torch": ty.dyn_emb_to_torch, "data_type_to_dtype": ty.data_type_to_dtype, "data_type_to_dyn_emb": ty.data_type_to_dyn_emb,
        "encode_meta_json_file_path": ck.encode_meta_json_file_path, "encode_checkpoint_file_path": ck.encode_checkpoint_file_path,
        "encode_counter_checkpoint_file_path": ck.encode_counter_checkpoint_file_path, "find_files": ck.find_files, "get_loading_files": ck.get_loading_files,
    }


def _params(fn):
    out = []
    for name, p in inspect.signature(fn).parameters.items():
        if name in ("self", "ctx"):
            continue
        prefix = "*" if p.kind == p.VAR_POSITIONAL else "**" if p.kind == p.VAR_KEYWORD else ""
        out.append((prefix + name, p.default))
    return out


def test_every_golden_signature_is_covered():
    gold = json.load(open(G))
    assert set(gold) == set(_ours())


@pytest.mark.parametrize("name", sorted(json.load(open(G))))
def test_parameters_match_reference(name):
    gold = json.load(open(G))[name]
    ours = _params(_ours()[name])
    want_names = [p["name"] for p in gold["params"]]
    assert [n for n, _ in ours[:len(want_names)]] == want_names, f"{name} ({gold['file']}:{gold['line']})"
    # parameters this package ADDS come after the reference's and have defaults (e.g. the `timestamp` of the table ops): every call
    # written against the reference binds identically
    for n, default in ours[len(want_names):]:
        assert default is not inspect.Parameter.empty or n.startswith("*"), f"{name}: extra parameter `{n}` needs a default"
    for (n, default), p in zip(ours, gold["params"]):
        d = p["default"]
        if d is None:
            continue                                    # required in the reference: a default here only widens what is accepted
        assert default is not inspect.Parameter.empty, f"{name}: `{n}` has a default in the reference"
        if "value" in d and n not in ("table_name",):      # literal default: must be equal.  Enum-member defaults (kept as source text) are checked below
            mine = list(default) if isinstance(default, tuple) else default          # the golden file is JSON: tuples are lists there
            assert mine == d["value"] and type(mine) is type(d["value"]), f"{name}: default of `{n}` is {default!r}, reference {d['value']!r}"


def test_enum_defaults_match_reference_members():
    """Defaults the reference writes as enum members (recorded as source text in the golden file)."""
    from dynamicemb import BatchedDynamicEmbeddingTablesV2 as M, BoundsCheckMode, DynamicEmbPoolingMode, EmbOptimType
    gold = {p["name"]: p["default"] for p in json.load(open(G))["BatchedDynamicEmbeddingTablesV2.__init__"]["params"] if p["default"]}
    sig = inspect.signature(M.__init__).parameters
    assert gold["pooling_mode"]["source"] == "DynamicEmbPoolingMode.SUM" and sig["pooling_mode"].default is DynamicEmbPoolingMode.SUM
    assert gold["bounds_check_mode"]["source"] == "BoundsCheckMode.WARNING" and sig["bounds_check_mode"].default is BoundsCheckMode.WARNING
    assert gold["optimizer"]["source"] == "EmbOptimType.SGD" and sig["optimizer"].default is EmbOptimType.SGD
    assert gold["output_dtype"]["source"] == "torch.float32"


# ---------------------------------------------------------------------------------------------------------------------------------
# configuration objects: dataclass fields (order, literal defaults) and enum members (order, literal values) — tests/golden/api_classes.json
GC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "api_classes.json")


def _our_class(name):
    import dynamicemb
    from dynamicemb import scored_hashtable, types
    for mod in (dynamicemb, types, scored_hashtable):
        if hasattr(mod, name):
            return getattr(mod, name)
    raise AssertionError(f"{name} is not offered by this package")


@pytest.mark.parametrize("name", sorted(json.load(open(GC))))
def test_config_classes_match_reference(name):
    import dataclasses
    import enum
    gold = json.load(open(GC))[name]
    cls = _our_class(name)
    if any("Enum" in b for b in gold["bases"]):
        assert issubclass(cls, enum.Enum) and (issubclass(cls, enum.IntEnum) == any("IntEnum" in b for b in gold["bases"]))
        assert [m.name for m in cls] == [f["name"] for f in gold["fields"]], f"{name} members ({gold['file']}:{gold['line']})"
        for m, f in zip(cls, gold["fields"]):
            if "value" in f["default"]:
                assert m.value == f["default"]["value"], f"{name}.{m.name}"
        return
    assert dataclasses.is_dataclass(cls), name
    ours = dataclasses.fields(cls)
    assert [f.name for f in ours] == [f["name"] for f in gold["fields"]], f"{name} fields ({gold['file']}:{gold['line']})"
    for mine, f in zip(ours, gold["fields"]):
        d = f["default"]
        if d is None:
            assert mine.default is dataclasses.MISSING and mine.default_factory is dataclasses.MISSING, f"{name}.{mine.name} is required in the reference"
        elif "value" in d:
            assert mine.default == d["value"] and type(mine.default) is type(d["value"]), f"{name}.{mine.name}: {mine.default!r} vs {d['value']!r}"
        else:
            assert mine.default is not dataclasses.MISSING or mine.default_factory is not dataclasses.MISSING, f"{name}.{mine.name} has a default in the reference"


def test_table_options_group_and_validate_like_the_reference():
    """dynamicemb_config.py:480-520: grouped key = 7 fields, options compare / hash by it; eval initializer must be constant; one-element
    score tuples unwrap, unsupported compounds raise."""
    from dynamicemb import DynamicEmbInitializerArgs, DynamicEmbInitializerMode, DynamicEmbScoreStrategy as S, DynamicEmbTableOptions as O
    a, b = O(dim=8, max_capacity=10, bucket_capacity=64), O(dim=128, max_capacity=99)
    assert list(a.get_grouped_key()) == ["training", "caching", "external_storage", "index_type", "dist_type", "score_strategy", "admit_strategy"]
    assert a == b and hash(a) == hash(b) and a != O(caching=True) and a != O(dist_type="continuous")
    assert O(score_strategy=(S.LFU,)).score_strategy is S.LFU and O(score_strategy=(S.LFU, S.TIMESTAMP)).score_strategy == (S.LFU, S.TIMESTAMP)
    for bad in ((S.LFU, S.STEP), (S.LFU, S.TIMESTAMP, S.LFU), ()):
        with pytest.raises(NotImplementedError):
            O(score_strategy=bad)
    with pytest.raises(TypeError):
        O(score_strategy=3)
    with pytest.raises(ValueError):
        O(dist_type="modulo")
    with pytest.raises(AssertionError):
        O(eval_initializer_args=DynamicEmbInitializerArgs(mode=DynamicEmbInitializerMode.UNIFORM))


def test_package_exports_every_name_of_the_reference():
    """`from dynamicemb import X` works for every X of the reference's `__all__` (tests/golden/api_exports.json)."""
    import dynamicemb
    names = json.load(open(os.path.join(os.path.dirname(G), "api_exports.json")))["dynamicemb.__all__"]
    assert len(names) >= 28
    missing = [n for n in names if not hasattr(dynamicemb, n)]
    assert not missing, missing


# import sites of the reference's user code (examples/, corelib/dynamicemb/example, benchmark): tests/golden/api_imports.json
_NOT_ON_THE_PATH = {
    "dynamicemb.exportable_tables": "inference export (torch.export of embedding collections) — SURVEY §2 out of scope",
}


def test_user_import_sites_of_the_reference_resolve():
    """Every `from dynamicemb[...] import name` the reference's examples and benchmarks contain resolves against this package, module path
    included, except the modules listed (with the reason) in _NOT_ON_THE_PATH."""
    import importlib
    sites = json.load(open(os.path.join(os.path.dirname(G), "api_imports.json")))
    assert set(_NOT_ON_THE_PATH) <= set(sites)
    checked = 0
    for module, names in sites.items():
        if module in _NOT_ON_THE_PATH:
            continue
        mod = importlib.import_module(module)
        for n in names:
            assert hasattr(mod, n), f"from {module} import {n}"
            checked += 1
    assert checked >= 28


def test_hstu_import_sites_of_the_reference_resolve():
    """examples/hstu imports `from hstu import hstu_attn_varlen_func`, `import hstu.hstu_ops_gpu` and, in its fused layer,
    `from hstu.hstu_blackwell import hstu_ops_gpu` (ops/fused_hstu_op.py:19-20, :50-58)."""
    import hstu
    import hstu.hstu_ops_gpu as ops
    from hstu import hstu_attn_varlen_func  # noqa: F401
    from hstu.hstu_blackwell import hstu_ops_gpu as bw
    assert bw.hstu_varlen_fwd_100 is ops.hstu_varlen_fwd_100 and bw.hstu_varlen_bwd_100 is ops.hstu_varlen_bwd_100 and hstu.hstu_ops_gpu is ops


def _callee(name):
    import dynamicemb
    from dynamicemb import dump_load as dl, planner, shard
    from hstu import hstu_attn_varlen_func, hstu_ops_gpu
    from hstu.fused_hstu_op import fused_hstu_op
    from dynamicemb.benchmark.dataset_generator import zipf
    from dynamicemb.get_planner import get_planner
    table = {"get_planner": get_planner, "gpu_zipf": zipf, "fused_hstu_op": fused_hstu_op, "hstu_varlen_fwd_100": hstu_ops_gpu.hstu_varlen_fwd_100, "hstu_varlen_bwd_100": hstu_ops_gpu.hstu_varlen_bwd_100,
             "dynamic_emb_save": dl.DynamicEmbDump, "dynamic_emb_load": dl.DynamicEmbLoad, "hstu_attn_varlen_func": hstu_attn_varlen_func,
             "incremental_dump": None, "get_score": None, "set_score": None}         # module-level AND method forms exist: checked separately
    if name in table:
        return table[name]
    for mod in (dynamicemb, planner, shard, dl):
        if hasattr(mod, name):
            return getattr(mod, name)
    raise AssertionError(name)


def test_call_shapes_of_the_reference_bind():
    """Every distinct (positional count, keyword set) with which the reference's examples, benchmarks and tests call a boundary class or
    function (tests/golden/api_call_sites.json) binds against this package's signature of the same name."""
    from dynamicemb import BatchedDynamicEmbeddingTablesV2 as M, LinearBucketTable as LT, dump_load as dl
    sites = json.load(open(os.path.join(os.path.dirname(G), "api_call_sites.json")))
    bound = 0
    for name, shapes in sites.items():
        target = _callee(name)
        for s in shapes:
            where = f"{name} at {s['file']}:{s['line']}"
            # three homonyms in the reference: function(model, ...), module.method(...), table.method(...)
            candidates = [target] if target is not None else [getattr(dl, name), getattr(M, name)] + ([getattr(LT, name)] if hasattr(LT, name) else [])
            errors = []
            for fn in candidates:
                sig = inspect.signature(fn)
                params = [p for p in sig.parameters.values() if p.name not in ("self",)]
                sig = sig.replace(parameters=params)
                try:
                    binder = sig.bind_partial if s["star_kwargs"] else sig.bind
                    binder(*([None] * s["positional"]), **{k: None for k in s["keywords"]})
                    break
                except TypeError as e:
                    errors.append(str(e))
            else:
                raise AssertionError(f"{where}: {errors}")
            bound += 1
    assert bound >= 60


def test_benchmark_id_generators():
    """dynamicemb.benchmark.dataset_generator (import site of examples/commons/datasets/hstu_batch.py:161): ranges, dtype and skew."""
    import torch
    from dynamicemb.benchmark.dataset_generator import PowerLaw, gen_key, translateToPowerLaw, zipf
    torch.manual_seed(0)
    ids = PowerLaw(1, 10 ** 9, 1.05, 200_000, device=torch.device("cpu"))
    assert ids.dtype == torch.int64 and int(ids.min()) >= 1 and int(ids.max()) < 10 ** 9
    assert float((ids < 1000).float().mean()) > 0.25 and float((ids > 10 ** 6).float().mean()) > 0.15      # heavy head AND a long tail
    x = torch.tensor([0.0, 0.5, 1.0 - 1e-12], dtype=torch.float64)
    y = translateToPowerLaw(1, 100, 1.05, x)
    assert float(y[0]) == 1.0 and 1.0 < float(y[1]) < 99.0 and float(y[2]) < 100.0 and int(y[2]) == 99
    assert gen_key(4, 3, 1.05, 1000, torch.device("cpu")).numel() == 12
    z = zipf(5, 105, 1.2, 50_000, torch.device("cpu"))
    assert z.dtype == torch.int64 and int(z.min()) >= 5 and int(z.max()) < 105
    counts = torch.bincount(z - 5, minlength=100).sort(descending=True).values.float()
    assert counts[0] / counts[9] > 8.0                                                                      # rank-1 vs rank-10: 10^1.2 = 15.8


def test_get_planner_builds_constraints_and_plans():
    """get_planner (reference get_planner.py:60-131): data-parallel / DynamicEmb / other tables get their constraints, the planner plans
    the DynamicEmb ones row-wise."""
    import torch
    from dynamicemb import DynamicEmbScoreStrategy, DynamicEmbTableOptions
    from dynamicemb.get_planner import get_planner
    from dynamicemb.utils import TORCHREC_TYPES

    class Cfg:
        def __init__(self, name, dim, num):
            self.name, self.embedding_dim, self.num_embeddings, self.feature_names = name, dim, num, [name]
    cfgs = [Cfg("item", 128, 1_000_000), Cfg("ctx", 64, 1000), Cfg("user", 128, 500_000)]
    planner = get_planner(cfgs, {"ctx"}, {"item": DynamicEmbTableOptions(score_strategy=DynamicEmbScoreStrategy.STEP)}, torch.device("cpu"))
    plan = planner.plan()
    assert set(plan) == {"item"} and plan["item"]["sharding_type"] == "row_wise" and plan["item"]["local_capacity"] == 1_000_064
    assert isinstance(TORCHREC_TYPES, set)
