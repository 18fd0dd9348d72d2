"""Golden API signatures: parameter names, order and literal defaults of the reference's boundary functions on the two hot paths, read from
the reference's source with `ast` (nothing is imported or executed).  Not run by the test suite; rerun by hand:
    python tests/golden/gen_golden_sigs.py        -> tests/golden/api_signatures.json
`tests/test_api_signatures_cpu.py` compares this package's functions of the same names against it."""
import ast
import json
import os

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "api_signatures.json")

TARGETS = {
    "examples/hstu/ops/fused_hstu_op.py": ["fused_hstu_op", "FusedHSTULayerFunction.forward"],
    "examples/hstu/ops/triton_ops/triton_layer_norm.py": ["triton_weighted_layer_norm_fwd", "triton_weighted_layer_norm_bwd"],
    "examples/hstu/ops/triton_ops/triton_norm_mul_dropout.py": ["triton_layer_norm_mul_dropout_fwd", "triton_layer_norm_mul_dropout_bwd"],
    "examples/hstu/ops/triton_ops/triton_silu.py": ["triton_silu_fwd", "triton_silu_bwd"],
    "third_party/FBGEMM/fbgemm_gpu/experimental/hstu/hstu/cuda_hstu_attention.py": ["hstu_attn_varlen_func"],
    "third_party/FBGEMM/fbgemm_gpu/experimental/hstu/src/hstu_blackwell/hstu_ops_gpu.py": ["hstu_varlen_fwd_100", "hstu_varlen_bwd_100"],
    "corelib/dynamicemb/dynamicemb/batched_dynamicemb_tables.py": [
        "BatchedDynamicEmbeddingTablesV2.__init__", "BatchedDynamicEmbeddingTablesV2.forward", "BatchedDynamicEmbeddingTablesV2.prefetch",
        "BatchedDynamicEmbeddingTablesV2.dump", "BatchedDynamicEmbeddingTablesV2.load", "BatchedDynamicEmbeddingTablesV2.export_keys_values",
        "BatchedDynamicEmbeddingTablesV2.set_score", "BatchedDynamicEmbeddingTablesV2.set_learning_rate",
        "BatchedDynamicEmbeddingTablesV2.incremental_dump", "BatchedDynamicEmbeddingTablesV2.fill_tables",
        "BatchedDynamicEmbeddingTablesV2.split_embedding_weights", "BatchedDynamicEmbeddingTablesV2.reset_cache_states",
        "BatchedDynamicEmbeddingTablesV2.set_record_cache_metrics", "BatchedDynamicEmbeddingTablesV2.flush", "BatchedDynamicEmbeddingTablesV2.get_score",
        "encode_meta_json_file_path", "encode_checkpoint_file_path", "encode_counter_checkpoint_file_path", "find_files", "get_loading_files"],
    "corelib/dynamicemb/dynamicemb/get_planner.py": ["get_planner"],
    "corelib/dynamicemb/benchmark/dataset_generator.py": ["translateToPowerLaw", "PowerLaw", "gen_key", "zipf"],
    "corelib/dynamicemb/dynamicemb/dump_load.py": ["find_sharded_modules", "get_dynamic_emb_module", "DynamicEmbDump", "DynamicEmbLoad"],
    "corelib/dynamicemb/dynamicemb/incremental_dump.py": ["set_score", "get_score", "incremental_dump", "is_valid_score_threshold"],
    "corelib/dynamicemb/dynamicemb/dynamicemb_config.py": ["get_sharded_table_capacity", "get_table_value_bytes", "string_to_evict_strategy",
                                                           "dyn_emb_to_torch", "data_type_to_dtype", "data_type_to_dyn_emb"],
    "corelib/dynamicemb/dynamicemb/embedding_admission.py": [
        "KVCounter.__init__", "MultiTableKVCounter.__init__", "MultiTableKVCounter.add", "MultiTableKVCounter.erase",
        "MultiTableKVCounter.memory_usage", "MultiTableKVCounter.load", "MultiTableKVCounter.dump",
        "FrequencyAdmissionStrategy.__init__", "FrequencyAdmissionStrategy.admit", "FrequencyAdmissionStrategy.initialize_non_admitted_embeddings"],
    "corelib/dynamicemb/dynamicemb/scored_hashtable.py": [
        "LinearBucketTable.__init__", "LinearBucketTable.lookup", "LinearBucketTable.insert", "LinearBucketTable.insert_and_evict",
        "LinearBucketTable.erase", "LinearBucketTable.load", "LinearBucketTable.dump", "get_scored_table"],
}


def literal(node):
    try:
        return {"value": ast.literal_eval(node)}
    except Exception:  # noqa: BLE001  (enum members, calls: keep the source text)
        return {"source": ast.unparse(node)}


def signature(fn: ast.FunctionDef):
    a = fn.args
    pos = [x.arg for x in a.posonlyargs + a.args]
    defaults = [None] * (len(pos) - len(a.defaults)) + [literal(d) for d in a.defaults]
    out = [{"name": n, "default": d} for n, d in zip(pos, defaults)]
    if a.vararg:
        out.append({"name": "*" + a.vararg.arg, "default": None})
    for x, d in zip(a.kwonlyargs, a.kw_defaults):
        out.append({"name": x.arg, "default": literal(d) if d is not None else None, "kwonly": True})
    if a.kwarg:
        out.append({"name": "**" + a.kwarg.arg, "default": None})
    return [p for p in out if p["name"] not in ("self", "ctx")]


def main():
    rec = {}
    for rel, names in TARGETS.items():
        tree = ast.parse(open(os.path.join(REF, rel)).read())
        index = {}
        for node in tree.body:
            if isinstance(node, ast.FunctionDef):
                index[node.name] = node
            elif isinstance(node, ast.ClassDef):
                for sub in node.body:
                    if isinstance(sub, ast.FunctionDef):
                        index[f"{node.name}.{sub.name}"] = sub
        for n in names:
            rec[n] = {"file": rel, "line": index[n].lineno, "params": signature(index[n])}
    json.dump(rec, open(OUT, "w"), indent=1)
    print(OUT, len(rec), "signatures")


if __name__ == "__main__":
    main()


# ---------------------------------------------------------------------------------------------------------------------------------
# dataclass fields (name, literal default) and enum members (name, literal value) of the configuration objects a caller constructs
CLASSES = {
    "corelib/dynamicemb/dynamicemb/dynamicemb_config.py": ["DynamicEmbTableOptions", "DynamicEmbInitializerArgs", "DynamicEmbInitializerMode",
                                                           "DynamicEmbCheckMode", "DynamicEmbPoolingMode", "DynamicEmbScoreStrategy",
                                                           "DynamicEmbEvictStrategy"],
    "corelib/dynamicemb/dynamicemb/types.py": ["DynamicEmbInitializerArgs", "DynamicEmbInitializerMode", "MemoryType"],
    "corelib/dynamicemb/dynamicemb/scored_hashtable.py": ["ScoreSpec", "ScoreArg", "ProbingType", "ReductionType"],
}
OUT_CLASSES = os.path.join(os.path.dirname(os.path.abspath(__file__)), "api_classes.json")


def main_classes():
    rec = {}
    for rel, names in CLASSES.items():
        tree = ast.parse(open(os.path.join(REF, rel)).read())
        for node in tree.body:
            if isinstance(node, ast.ClassDef) and node.name in names and node.name not in rec:
                fields = []
                for sub in node.body:
                    if isinstance(sub, ast.AnnAssign) and isinstance(sub.target, ast.Name):
                        fields.append({"name": sub.target.id, "default": literal(sub.value) if sub.value is not None else None})
                    elif isinstance(sub, ast.Assign) and len(sub.targets) == 1 and isinstance(sub.targets[0], ast.Name):
                        fields.append({"name": sub.targets[0].id, "default": literal(sub.value)})
                rec[node.name] = {"file": rel, "line": node.lineno, "bases": [ast.unparse(b) for b in node.bases], "fields": fields}
    json.dump(rec, open(OUT_CLASSES, "w"), indent=1)
    print(OUT_CLASSES, sorted(rec))


if __name__ == "__main__":
    main_classes()


# ---------------------------------------------------------------------------------------------------------------------------------
# names the reference package exports (`dynamicemb.__all__`) and the model-level checkpoint / score functions
OUT_EXPORTS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "api_exports.json")


def main_exports():
    tree = ast.parse(open(os.path.join(REF, "corelib/dynamicemb/dynamicemb/__init__.py")).read())
    names = []
    for node in ast.walk(tree):
        if isinstance(node, ast.Assign) and isinstance(node.targets[0], ast.Name) and node.targets[0].id == "__all__":
            names = ast.literal_eval(node.value)
    json.dump({"dynamicemb.__all__": sorted(names)}, open(OUT_EXPORTS, "w"), indent=1)
    print(OUT_EXPORTS, len(names), "names")


if __name__ == "__main__":
    main_exports()


# ---------------------------------------------------------------------------------------------------------------------------------
# import sites: what the reference's USER code (examples/, corelib/dynamicemb/example, corelib/dynamicemb/benchmark) imports from dynamicemb
OUT_IMPORTS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "api_imports.json")


def main_imports():
    sites = {}
    for top in ("examples", "corelib/dynamicemb/example", "corelib/dynamicemb/benchmark"):
        for dp, _, fs in os.walk(os.path.join(REF, top)):
            for f in fs:
                if not f.endswith(".py"):
                    continue
                try:
                    tree = ast.parse(open(os.path.join(dp, f)).read())
                except SyntaxError:
                    continue
                for node in ast.walk(tree):
                    if isinstance(node, ast.ImportFrom) and node.module and (node.module == "dynamicemb" or node.module.startswith("dynamicemb.")
                                                                             or node.module == "dynamicemb_extensions"):
                        for a in node.names:
                            sites.setdefault(node.module, set()).add(a.name)
    json.dump({k: sorted(v) for k, v in sorted(sites.items())}, open(OUT_IMPORTS, "w"), indent=1)
    print(OUT_IMPORTS, {k: len(v) for k, v in sites.items()})


if __name__ == "__main__":
    main_imports()


# ---------------------------------------------------------------------------------------------------------------------------------
# call sites: keyword arguments (and positional counts) the reference's user code passes to the boundary classes / functions
OUT_CALLS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "api_call_sites.json")
CALLEES = ["DynamicEmbTableOptions", "DynamicEmbInitializerArgs", "DynamicEmbParameterConstraints", "DynamicEmbeddingShardingPlanner",
           "DynamicEmbeddingEnumerator", "DynamicEmbeddingCollectionSharder", "DynamicEmbeddingBagCollectionSharder", "FrequencyAdmissionStrategy",
           "KVCounter", "BatchedDynamicEmbeddingTablesV2", "DynamicEmbDump", "DynamicEmbLoad", "dynamic_emb_save", "dynamic_emb_load",
           "get_sharded_table_capacity", "get_table_value_bytes", "hstu_attn_varlen_func", "incremental_dump", "get_score", "set_score",
           "fused_hstu_op", "hstu_varlen_fwd_100", "hstu_varlen_bwd_100", "get_planner", "gpu_zipf"]


def main_calls():
    sites = {}
    for top in ("examples", "corelib/dynamicemb/example", "corelib/dynamicemb/benchmark", "corelib/dynamicemb/test"):
        for dp, _, fs in os.walk(os.path.join(REF, top)):
            for f in fs:
                if not f.endswith(".py"):
                    continue
                try:
                    tree = ast.parse(open(os.path.join(dp, f)).read())
                except SyntaxError:
                    continue
                local_defs = {n.name for n in ast.walk(tree) if isinstance(n, (ast.FunctionDef, ast.ClassDef))}      # homonyms defined in the file itself
                for node in ast.walk(tree):
                    if not isinstance(node, ast.Call):
                        continue
                    fn = node.func
                    name = fn.id if isinstance(fn, ast.Name) else fn.attr if isinstance(fn, ast.Attribute) else None
                    if name not in CALLEES or name in local_defs or any(isinstance(a, ast.Starred) for a in node.args):
                        continue
                    kws = sorted(k.arg for k in node.keywords if k.arg is not None)
                    rec = {"file": os.path.relpath(os.path.join(dp, f), REF), "line": node.lineno, "positional": len(node.args), "keywords": kws,
                           "star_kwargs": any(k.arg is None for k in node.keywords)}
                    key = (rec["positional"], tuple(kws))
                    sites.setdefault(name, {}).setdefault(str(key), rec)          # one example per distinct call shape
    out = {k: list(v.values()) for k, v in sorted(sites.items())}
    json.dump(out, open(OUT_CALLS, "w"), indent=1)
    print(OUT_CALLS, {k: len(v) for k, v in out.items()})


if __name__ == "__main__":
    main_calls()
