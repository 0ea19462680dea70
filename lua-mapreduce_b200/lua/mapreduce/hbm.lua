-- mapreduce/hbm.lua -- the "hbm" storage of lua-mapreduce: replaces the text spill through
-- GridFS / shared FS / scp (mapreduce/fs.lua:185-208) by the in-HBM shuffle of the mrhbm C
-- module.  The seam sits one level above fs.router: job.lua's emit closure (job.lua:83-97) and
-- the consumer of utils.merge_iterator (job.lua:264-284) call this module, so no "%q" line is
-- ever formatted or load()-ed.
--
-- NOT EXECUTABLE IN THE BUILD IMAGE (no Lua).  The Python mirror in ../../mapreduce/ runs
-- the same call sequence against the same C ABI under test.
local mrhbm = require "mrhbm"

local hbm = { _VERSION = "0.1", _NAME = "mapreduce.hbm" }

-- what the device evaluates is declared on the plugin tables, like the reducer algebra flags
-- (examples/WordCount/reducefn.lua:10-14):
--   partitionfn module: hbm_partitionfn = "fnv_lua" | "wordhash" | "mulhash", NUM_REDUCERS = n
--   reducefn / combinerfn module: hbm_reducefn = "sum"
local ctx -- one context per worker process (job.lua keeps `funcs`/`initialized` per process too)

function hbm.configure(partition_mod, reduce_mod, combiner_mod, opts)
  -- hbm_reducefn = "sum": reduced on the device.  No declaration: general reducer -- the device
  -- partitions, sorts and groups, job.lua:264-284 calls reducefn per group on the host.
  local builtin = reduce_mod.hbm_reducefn == "sum"
  assert(not combiner_mod or (builtin and combiner_mod.hbm_reducefn == "sum"),
         "a combinerfn runs on the device: it (and the reducefn) must declare hbm_reducefn = 'sum'")
  opts = opts or {}
  local c, err = mrhbm.new{
    key_kind = opts.key_kind or "str",
    max_key_bytes = opts.max_key_bytes or 123,
    num_partitions = assert(partition_mod.NUM_REDUCERS, "partitionfn module must expose NUM_REDUCERS"),
    partitioner = assert(partition_mod.hbm_partitionfn, "partitionfn module must declare hbm_partitionfn"),
    combiner = combiner_mod ~= nil,
    reducer = builtin and "sum" or "none",
    device = opts.device,
  }
  ctx = assert(c, err)
  return ctx
end

-- job_prepare_map (job.lua:154-228): returns the emit closure and the finisher
function hbm.map_job(map_key)
  local m = assert(ctx:map_begin(tostring(map_key)))
  local emit = function(key, value) assert(m:emit(key, value)) end
  local finish = function(ok)
    if ok then assert(m:commit()) else m:abort() end -- BROKEN jobs publish nothing
  end
  return emit, finish
end

-- server_prepare_reduce (server.lua:279-329)
function hbm.prepare_reduce()
  assert(ctx:shuffle())
  return assert(ctx:partitions())
end

-- job_prepare_reduce (job.lua:230-296): for k,v in hbm.groups(part_key) do ... end
function hbm.groups(part_key) return assert(ctx:groups(part_key)) end

function hbm.reset() assert(ctx:reset()) end

return hbm
