"""Call-site compatibility, checked statically against the reference tree (CPU; skipped where /root/reference is absent).

Every call the reference's own Python code makes into its binding module (`wmb.<name>(...)` inside pylibwholegraph/torch/*.py
and inside its tests) and every call its tests make into the torch layer (`wgth.<name>(...)`, `wg_ops.<name>(...)`,
`graph_ops.<name>(...)`) is located with `ast`, and the same call shape -- number of positional arguments, keyword names --
must bind to this repo's function of the same name (`inspect.signature(...).bind`).  Calls with *args / **kwargs are skipped.
This is what lets the reference's torch layer and tests drive this implementation without edits."""
import ast
import glob
import importlib
import inspect
import os

import pytest

REF_PY = "/root/reference/python/pylibwholegraph/pylibwholegraph"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF_PY), reason="reference tree not present")


OURS = {"pylibwholegraph.binding.wholememory_binding": "wholegraph_b200.binding", "pylibwholegraph.torch": "wholegraph_b200.torch",
        "pylibwholegraph.utils": "wholegraph_b200.utils", "pylibwholegraph.test_utils": "wholegraph_b200.test_utils"}
# the reference's GNN example glue and launch helpers: outside SURVEY.md section 8
OUT_OF_SCOPE_FILES = {"gnn_model.py", "data_loader.py", "common_options.py", "distributed_launch.py", "gat_conv.py", "sage_conv.py"}
# call sites that are wrong in the reference itself: embedding.py:336,339 call get_stream(False), but its get_stream()
# (wholegraph_env.py:27) takes no argument -- writeback_all_cache / drop_all_cache raise TypeError there
REFERENCE_BUGS = {("torch/embedding.py", "get_stream")}
OUT_OF_SCOPE_NAMES = {"distributed_launch", "add_distributed_launch_options", "get_rank", "get_world_size", "get_local_rank", "get_local_size"}


def _our_module(ref_name):
    """pylibwholegraph.torch[.x] / pylibwholegraph.binding.wholememory_binding -> this repo's module, else None"""
    for ref, ours in OURS.items():
        if ref_name == ref or ref_name.startswith(ref + "."):
            try:
                return importlib.import_module(ours + ref_name[len(ref):])
            except ModuleNotFoundError:
                return None
    return None


def _aliases(tree, path):
    """local name -> this repo's module, from the file's import statements (absolute, and relative inside torch/)"""
    package = "pylibwholegraph.torch" if os.path.dirname(path) == REF_PY + "/torch" else None
    out = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.Import):
            for a in node.names:
                if a.asname:
                    m = _our_module(a.name)
                    if m is not None:
                        out[a.asname] = m
        elif isinstance(node, ast.ImportFrom):
            base = node.module or ""
            if node.level == 1 and package:
                base = package + ("." + base if base else "")
            elif node.level == 2 and package:
                base = "pylibwholegraph" + ("." + base if base else "")
            elif node.level:
                continue
            for a in node.names:
                m = _our_module(base + "." + a.name)
                if m is not None:
                    out[a.asname or a.name] = m
                    continue
                parent = _our_module(base)  # `from <module> import <function or class>`
                if parent is not None and a.name != "*":
                    out[a.asname or a.name] = getattr(parent, a.name, _Missing(base + "." + a.name))
    return out


class _Missing(object):
    def __init__(self, what):
        self.what = what


def _files():
    files = sorted(glob.glob(REF_PY + "/torch/*.py") + glob.glob(REF_PY + "/tests/**/*.py", recursive=True)
                   + glob.glob(REF_PY + "/test_utils/*.py"))
    return [f for f in files if os.path.basename(f) not in OUT_OF_SCOPE_FILES]


def _call_sites():
    for f in _files():
        tree = ast.parse(open(f).read(), f)
        aliases = _aliases(tree, f)
        for node in ast.walk(tree):
            if isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute) and isinstance(node.func.value, ast.Name) \
                    and inspect.ismodule(aliases.get(node.func.value.id)):
                yield f, node.lineno, node.func.value.id, aliases[node.func.value.id], node.func.attr, node
            elif isinstance(node, ast.Call) and isinstance(node.func, ast.Name) and node.func.id in aliases \
                    and not inspect.ismodule(aliases[node.func.id]):
                yield f, node.lineno, "", None, node.func.id, node  # a name imported with `from ... import`


def test_every_reference_call_site_binds_to_this_implementation():
    checked, problems = 0, []
    for f, line, base, module, name, node in _call_sites():
        if name in OUT_OF_SCOPE_NAMES or (os.path.relpath(f, REF_PY), name) in REFERENCE_BUGS:
            continue
        where = "%s:%d %s.%s" % (os.path.relpath(f, REF_PY), line, base, name)
        if module is None:
            fn = _aliases(ast.parse(open(f).read(), f), f)[name]
            fn = None if isinstance(fn, _Missing) else fn
        else:
            fn = getattr(module, name, None)
        if fn is None:
            problems.append(where + ": missing")
            continue
        if any(isinstance(a, ast.Starred) for a in node.args) or any(k.arg is None for k in node.keywords):
            continue
        if inspect.isclass(fn):
            sig_of = fn.__init__
            args = [None] * (len(node.args) + 1)  # self
        else:
            sig_of, args = fn, [None] * len(node.args)
        try:
            sig = inspect.signature(sig_of)
        except (TypeError, ValueError):
            continue  # enum classes etc.
        try:
            sig.bind(*args, **{k.arg: None for k in node.keywords})
            checked += 1
        except TypeError as e:
            problems.append("%s: %s (ours: %s)" % (where, e, sig))
    assert not problems, "\n".join(problems)
    assert checked > 100, checked


def test_attribute_reads_on_aliased_modules_exist():
    """`wmb.<Name>` / `wgth.<name>` reads that are not calls (enum classes, constants, functions passed around) must exist too."""
    missing = set()
    for f in _files():
        tree = ast.parse(open(f).read(), f)
        aliases = _aliases(tree, f)
        for node in ast.walk(tree):
            if isinstance(node, ast.Attribute) and isinstance(node.value, ast.Name) and inspect.ismodule(aliases.get(node.value.id)) \
                    and node.attr not in OUT_OF_SCOPE_NAMES and not hasattr(aliases[node.value.id], node.attr):
                missing.add("%s:%d %s.%s" % (os.path.relpath(f, REF_PY), node.lineno, node.value.id, node.attr))
    assert not missing, sorted(missing)


def _our_methods():
    import wholegraph_b200.binding as wmb
    import wholegraph_b200.torch as wgth
    import wholegraph_b200.torch.wholegraph_env as env
    methods = {}
    for mod in (wmb, wgth, env):
        for cname, cls in inspect.getmembers(mod, inspect.isclass):
            if cls.__module__.startswith("wholegraph_b200"):
                for mname, fn in inspect.getmembers(cls, inspect.isfunction):
                    if not mname.startswith("__"):
                        methods.setdefault(mname, {})[cname] = fn
    return methods


# method names that belong to something else at that call site (torch.Tensor.scatter / torch.gather on plain tensors in the
# reference's op test, torch.utils.cpp_extension.load)
NOT_OUR_OBJECT = {("tests/wholegraph_torch/ops/test_wholegraph_gather_scatter.py", "scatter"),
                  ("tests/wholegraph_torch/ops/test_wholegraph_gather_scatter.py", "gather"), ("torch/wholegraph_env.py", "load")}


def test_method_calls_bind_to_a_class_of_this_implementation():
    """`obj.method(...)` calls whose method name one of this repo's binding / torch classes defines: the call shape must bind
    to at least one of them (the receiver's type is not known statically)."""
    methods = _our_methods()
    checked, problems = 0, []
    for f in _files():
        tree = ast.parse(open(f).read(), f)
        aliases = _aliases(tree, f)
        rel = os.path.relpath(f, REF_PY)
        for node in ast.walk(tree):
            if not (isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute) and node.func.attr in methods):
                continue
            if isinstance(node.func.value, ast.Name) and node.func.value.id in aliases:
                continue  # module-level call, covered above
            if (rel, node.func.attr) in NOT_OUR_OBJECT:
                continue
            if any(isinstance(a, ast.Starred) for a in node.args) or any(k.arg is None for k in node.keywords):
                continue
            checked += 1
            for fn in methods[node.func.attr].values():
                try:
                    inspect.signature(fn).bind(None, *[None] * len(node.args), **{k.arg: None for k in node.keywords})
                    break
                except TypeError:
                    pass
            else:
                problems.append("%s:%d .%s(%d positional, keywords %s)" % (rel, node.lineno, node.func.attr, len(node.args),
                                                                          [k.arg for k in node.keywords]))
    assert not problems, "\n".join(problems)
    assert checked > 200, checked
