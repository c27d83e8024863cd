"""Import-path compatibility: `import pylibwholegraph.torch as wgth` and
`import pylibwholegraph.binding.wholememory_binding as wmb` resolve to this repo's implementation of the same API
(wholegraph_b200.torch / wholegraph_b200.binding).  Put <repo>/compat on PYTHONPATH next to <repo>; see INTEGRATION.md.

The helper modules the reference's tests import (`pylibwholegraph.utils.multiprocess`, `pylibwholegraph.test_utils.test_comm`)
resolve too.  Only the modules on the WholeMemory hot path exist: the GNN example glue of the reference package
(gnn_model, data_loader, common_options, distributed_launch, cugraphops) is out of scope and importing it raises
ModuleNotFoundError."""

# what the reference package exposes at its top level (pylibwholegraph/__init__.py): the API level this implementation
# mirrors, and no git commit (that is only non-empty in a built distribution of the reference)
__version__ = "24.12.00"
__git_commit__ = ""
__all__ = ["__git_commit__", "__version__"]
