--[[ dcgansr.lua -- LuaJIT FFI shim over libdcgansr.so (include/dcgansr.h).

Re-creates the Torch7 surface the reference scripts use (train.lua:97-283) on top of the C ABI:
   nn.Sequential():add(nn.SpatialFullConvolution(...))  ->  dcgansr.nn.Sequential():add(dcgansr.nn.SpatialFullConvolution(...))
   net:cuda()                                            ->  net:cuda(ctx, {nc, h, w}, batchSize)
   net:forward / :backward / :updateGradInput / :getParameters / :zeroGradParameters
   optim.adam(feval, parameters, optimState)             ->  dcgansr.optim.adam(net, optimState)   (update on the device vectors)
and adds the fused step   dcgansr.train_step(ctx, netG, netD, stepcfg, real_FloatTensor)  (fDx -> adam(D) -> fGx -> adam(G)).

NOTE: this file could not be executed where it was written (no Lua / LuaJIT / Torch7 in the build image); the
same header is exercised through Python ctypes/cffi by tests/test_abi.py and the GPU suite.  Host tensors are
torch.FloatTensor (NCHW, contiguous): only :data() and :size() are used.  ]]
local ffi = require 'ffi'

local M = {nn = {}, optim = {}}

local function read_header()
   local dir = os.getenv('DCGANSR_HOME') or '.'
   local f = assert(io.open(dir .. '/include/dcgansr.h', 'r'))
   local src = f:read('*a'); f:close()
   -- the declarations between the two markers are plain C99 (no macros, no preprocessor lines)
   local b = select(2, src:find('FFI%-CDEF%-BEGIN[^\n]*\n'))
   local e = src:find('/%* FFI%-CDEF%-END')
   return src:sub(b + 1, e - 1)
end
ffi.cdef(read_header())
local lib = ffi.load((os.getenv('DCGANSR_HOME') or '.') .. '/dcgan_super_resolution_b200/libdcgansr.so')
M.lib = lib

local function check(rc, ctx)
   if rc ~= 0 then error('libdcgansr error ' .. rc .. ': ' .. ffi.string(lib.dcgansr_last_error(ctx)), 2) end
end

-- ---------------------------------------------------------------- context (require 'cunn'; cutorch.setDevice)
function M.Context(opt)
   opt = opt or {}
   local cfg = ffi.new('dcgansr_cfg', {device = (opt.gpu or 1) - 1, precision = opt.precision == 'strict' and 0 or 1,
                                      world_size = opt.world_size or 1, rank = opt.rank or 0,
                                      sync_bn = opt.sync_bn and 1 or 0, use_graph = opt.use_graph and 1 or 0})
   local out = ffi.new('dcgansr_ctx*[1]')
   check(lib.dcgansr_ctx_create(cfg, out), nil)
   -- finalizer order between a ctx and its nets is not defined in LuaJIT: dcgansr_ctx_destroy releases the device memory of
   -- every net still alive and leaves plan-only handles behind, so either order is safe
   return ffi.gc(out[0], lib.dcgansr_ctx_destroy)
end

-- ---------------------------------------------------------------- module descriptors (torch/nn constructors)
local KIND = {conv = 1, fullconv = 2, bn = 3, relu = 4, lrelu = 5, tanh = 6, sigmoid = 7, upnearest = 8, view = 9}
local function layer(t) return t end
function M.nn.SpatialConvolution(nIn, nOut, kW, kH, dW, dH, padW, padH)
   return layer{kind = KIND.conv, cin = nIn, cout = nOut, kh = kH, kw = kW, sh = dH or 1, sw = dW or 1, ph = padH or 0, pw = padW or 0}
end
function M.nn.SpatialFullConvolution(nIn, nOut, kW, kH, dW, dH, padW, padH, adjW, adjH)
   return layer{kind = KIND.fullconv, cin = nIn, cout = nOut, kh = kH, kw = kW, sh = dH or 1, sw = dW or 1, ph = padH or 0,
                pw = padW or 0, adjh = adjH or 0, adjw = adjW or 0}
end
function M.nn.SpatialBatchNormalization(n, eps, momentum)
   return layer{kind = KIND.bn, cin = n, cout = n, eps = eps or 1e-5, momentum = momentum or 0.1}
end
function M.nn.ReLU() return layer{kind = KIND.relu} end
function M.nn.LeakyReLU(negval) return layer{kind = KIND.lrelu, negval = negval or 0.01} end
function M.nn.Tanh() return layer{kind = KIND.tanh} end
function M.nn.Sigmoid() return layer{kind = KIND.sigmoid} end
function M.nn.SpatialUpSamplingNearest(scale) return layer{kind = KIND.upnearest, scale = scale} end
function M.nn.View() local v = layer{kind = KIND.view}; v.setNumInputDims = function(self) return self end; return v end

-- ---------------------------------------------------------------- nn.Sequential
local Sequential = {}
Sequential.__index = Sequential
function M.nn.Sequential() return setmetatable({layers = {}}, Sequential) end
function Sequential:add(m) table.insert(self.layers, m); return self end
function Sequential:apply(fn) return self end        -- weights_init runs on the host: see weights_init below
function Sequential:cuda(ctx, in_shape, max_batch)
   local n = #self.layers
   local arr = ffi.new('dcgansr_layer[?]', n)
   for i, l in ipairs(self.layers) do
      for k, v in pairs(l) do if type(v) == 'number' then arr[i - 1][k] = v end end
   end
   local out = ffi.new('dcgansr_net*[1]')
   check(lib.dcgansr_net_create(ctx, arr, n, in_shape[1], in_shape[2], in_shape[3], max_batch, out), ctx)
   self.ctx, self.h, self.in_shape = ctx, ffi.gc(out[0], lib.dcgansr_net_destroy), in_shape
   local c, h, w = ffi.new('int[1]'), ffi.new('int[1]'), ffi.new('int[1]')
   check(lib.dcgansr_net_out_shape(self.h, c, h, w), ctx)
   self.out_shape = {c[0], h[0], w[0]}
   local np = ffi.new('int64_t[1]'); check(lib.dcgansr_net_num_params(self.h, np), ctx)
   self.nparams = tonumber(np[0])
   return self
end
function Sequential:forward(x)           -- x: torch.FloatTensor B x C x H x W
   local B = x:size(1)
   self.output = self.output or torch.FloatTensor()
   self.output:resize(B, self.out_shape[1], self.out_shape[2], self.out_shape[3])
   check(lib.dcgansr_net_forward(self.h, x:data(), B, self.output:data()), self.ctx)
   return self.output
end
function Sequential:backward(x, dy)
   self.gradInput = self.gradInput or torch.FloatTensor()
   self.gradInput:resizeAs(x)
   check(lib.dcgansr_net_backward(self.h, x:data(), dy:data(), x:size(1), self.gradInput:data()), self.ctx)
   return self.gradInput
end
function Sequential:updateGradInput(x, dy)
   self.gradInput = self.gradInput or torch.FloatTensor()
   self.gradInput:resizeAs(x)
   check(lib.dcgansr_net_update_grad_input(self.h, x:data(), dy:data(), x:size(1), self.gradInput:data()), self.ctx)
   return self.gradInput
end
function Sequential:zeroGradParameters() check(lib.dcgansr_net_zero_grads(self.h), self.ctx) end
-- getParameters(): host COPIES of the flat vectors (the live ones stay on the device)
function Sequential:getParameters()
   local p, g = torch.FloatTensor(self.nparams), torch.FloatTensor(self.nparams)
   check(lib.dcgansr_net_get_params(self.h, p:data()), self.ctx)
   check(lib.dcgansr_net_get_grads(self.h, g:data()), self.ctx)
   return p, g
end
function Sequential:setParameters(p) check(lib.dcgansr_net_set_params(self.h, p:data()), self.ctx) end

-- weights_init (train.lua:42-51) on the flat vector: conv ~ N(0, .02), BN gamma ~ N(1, .02), beta = 0
function M.weights_init(net)
   local p = torch.FloatTensor(net.nparams)
   local off = 1
   for _, l in ipairs(net.layers) do
      if l.kind == KIND.conv or l.kind == KIND.fullconv then
         local n = l.cin * l.cout * l.kh * l.kw
         p:narrow(1, off, n):normal(0.0, 0.02); off = off + n
      elseif l.kind == KIND.bn then
         p:narrow(1, off, l.cin):normal(1.0, 0.02); off = off + l.cin
         p:narrow(1, off, l.cin):fill(0); off = off + l.cin
      end
   end
   net:setParameters(p)
end

-- optim.adam(feval, x, state): the closure has already run (forward/backward above); update on the device
function M.optim.adam(net, state)
   check(lib.dcgansr_net_adam(net.h, state.learningRate or 1e-3, state.beta1 or 0.9, state.beta2 or 0.999, state.epsilon or 1e-8), net.ctx)
end

-- ---------------------------------------------------------------- the fused step
function M.StepCfg(t)
   return ffi.new('dcgansr_step_cfg', {loss = t.criterion == 'MSE' and 1 or 0, real_label = t.real_label, fake_label = t.fake_label or 0,
                                      gen_label = t.gen_label, pixel_label = t.pixel_label and 1 or 0, pixel_div = t.pixel_div or 1,
                                      lr = t.lr or 2e-4, beta1 = t.beta1 or 0.5, beta2 = t.beta2 or 0.999, eps = t.epsilon or 1e-8})
end
local losses = ffi.new('float[3]')
function M.train_step(ctx, netG, netD, cfg, real)     -- returns errD_real, errD_fake, errG
   check(lib.dcgansr_train_step(ctx, netG.h, netD.h, cfg, real:data(), real:size(1), losses), ctx)
   return losses[0], losses[1], losses[2]
end
-- the same step on a batch already resident on the device (dcgansr.stage_batch / dcgansr.stage_patches filled `slot`)
function M.stage_batch(ctx, netD, real, slot)
   check(lib.dcgansr_stage_batch(ctx, netD.h, real:data(), real:size(1), slot or 0), ctx)
end
function M.train_step_staged(ctx, netG, netD, cfg, slot, batch)
   check(lib.dcgansr_train_step_staged(ctx, netG.h, netD.h, cfg, slot or 0, batch, losses), ctx)
   return losses[0], losses[1], losses[2]
end

-- ---------------------------------------------------------------- patches, eval metrics, stitching (SURVEY 8(f))
-- real_none of the patch scripts without the per-pixel loops (train-gray-patch.lua:267-275): images K x H x W FloatTensor
function M.stage_patches(ctx, netD, images, patchSize, line, nper, stride, slot)
   check(lib.dcgansr_stage_patches(ctx, netD.h, images:data(), images:size(1), images:size(2), images:size(3), patchSize, line, nper,
                                   stride, slot or 0), ctx)
   return images:size(1) * nper
end
function M.extract_patches(ctx, images, patchSize, line, nper, stride)
   local out = torch.FloatTensor(images:size(1) * nper, patchSize, patchSize)
   check(lib.dcgansr_extract_patches(ctx, images:data(), out:data(), images:size(1), images:size(2), images:size(3), patchSize, line,
                                     nper, stride), ctx)
   return out
end
function M.assemble_patches(ctx, patches, images, patchSize, line, nper, stride)     -- images: K x H x W, overwritten where covered
   check(lib.dcgansr_assemble_patches(ctx, patches:data(), images:data(), images:size(1), images:size(2), images:size(3), patchSize,
                                      line, nper, stride), ctx)
   return images
end
-- calPSNR / calSSIM (train-gray-3.lua:143-221) on n x H x W pairs -> FloatTensor(n)
local function metric(fn, ctx, a, b)
   local out = torch.FloatTensor(a:size(1))
   check(fn(ctx, a:data(), b:data(), out:data(), a:size(1), a:size(2), a:size(3)), ctx)
   return out
end
function M.calPSNR(ctx, a, b) return metric(lib.dcgansr_psnr, ctx, a, b) end
function M.calSSIM(ctx, a, b) return metric(lib.dcgansr_ssim, ctx, a, b) end
-- image.scale(src, W, H) bilinear baseline (train-gray-3.lua:399) on n x h x w -> n x H x W
function M.scale_bilinear(ctx, src, H, W)
   local out = torch.FloatTensor(src:size(1), H, W)
   check(lib.dcgansr_scale_bilinear(ctx, src:data(), out:data(), src:size(1), src:size(2), src:size(3), H, W), ctx)
   return out
end
-- minimum-error boundary cut stitching (train-gray-patch-batch-overlap.lua:457-694): patches (k*L*L) x p x p -> k x fine x fine
function M.stitch_overlap(ctx, patches, fineSize, patchSize, overlap, k, fix_top_cost)
   local out = torch.FloatTensor(k or 1, fineSize, fineSize):zero()
   check(lib.dcgansr_stitch_overlap(ctx, patches:data(), out:data(), k or 1, fineSize, fineSize, patchSize, overlap,
                                    fix_top_cost and 1 or 0), ctx)
   return out
end

return M
