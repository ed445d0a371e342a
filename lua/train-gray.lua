--[[ train-gray.lua on libdcgansr.so: the host keeps the reference's opt / netG / netD surface
(/root/reference/train-gray.lua:9-157, 205-301); the hot path is one dcgansr.train_step per iteration.
Not executed in the build image (no LuaJIT/Torch7 there); see INTEGRATION.md. ]]
require 'torch'
local dsr = require 'dcgansr'
local nn = dsr.nn

opt = {dataset = 'folder', batchSize = 64, loadSize = 64, fineSize = 64, ngf = 16, ndf = 64, nThreads = 4, niter = 25,
       lr = 0.0002, beta1 = 0.5, ntrain = math.huge, gpu = 1, name = 'experiment1', precision = 'tf32'}
for k, v in pairs(opt) do opt[k] = tonumber(os.getenv(k)) or os.getenv(k) or opt[k] end      -- train-gray.lua:24
print(opt)
torch.manualSeed(torch.random(1, 10000))
torch.setdefaulttensortype('torch.FloatTensor')

local DataLoader = paths.dofile('data/data.lua')          -- the reference's own loader, unchanged
local data = DataLoader.new(opt.nThreads, opt.dataset, opt)

local nc, ndf, ngf = 1, opt.ndf, opt.ngf
local ctx = dsr.Context{gpu = opt.gpu, precision = opt.precision}

local netG = nn.Sequential()                                                  -- train-gray.lua:102-117
netG:add(nn.SpatialUpSamplingNearest(2))
netG:add(nn.SpatialFullConvolution(nc, ngf * 4, 4, 4, 2, 2, 1, 1))
netG:add(nn.SpatialFullConvolution(ngf * 4, ngf * 2, 4, 4, 2, 2, 1, 1))
netG:add(nn.SpatialConvolution(ngf * 2, ngf, 4, 4, 2, 2, 1, 1))
netG:add(nn.SpatialConvolution(ngf, nc, 4, 4, 2, 2, 1, 1))
netG:add(nn.Tanh())

local netD = nn.Sequential()                                                  -- train-gray.lua:121-137
netD:add(nn.SpatialConvolution(nc, ndf, 4, 4, 2, 2, 1, 1)):add(nn.LeakyReLU(0.2, true))
netD:add(nn.SpatialConvolution(ndf, ndf * 2, 4, 4, 2, 2, 1, 1)):add(nn.SpatialBatchNormalization(ndf * 2)):add(nn.LeakyReLU(0.2, true))
netD:add(nn.SpatialConvolution(ndf * 2, ndf * 4, 4, 4, 2, 2, 1, 1)):add(nn.SpatialBatchNormalization(ndf * 4)):add(nn.LeakyReLU(0.2, true))
netD:add(nn.SpatialConvolution(ndf * 4, ndf * 8, 4, 4, 2, 2, 1, 1)):add(nn.SpatialBatchNormalization(ndf * 8)):add(nn.LeakyReLU(0.2, true))
netD:add(nn.SpatialConvolution(ndf * 8, 1, 4, 4)):add(nn.Sigmoid()):add(nn.View(1):setNumInputDims(3))

netG:cuda(ctx, {nc, opt.fineSize / 2, opt.fineSize / 2}, opt.batchSize)      -- replaces :cuda() / cudnn.convert (:168-180)
netD:cuda(ctx, {nc, opt.fineSize, opt.fineSize}, 2 * opt.batchSize)              -- 2B samples: D(real) and D(fake) run as one grouped pass
dsr.weights_init(netG); dsr.weights_init(netD)                               -- netG:apply(weights_init) (:42-51)

-- criterion = nn.MSECriterion(); labels .001 / per-sample pixel MSE / 0  (train-gray.lua:140,232,265,282)
local step = dsr.StepCfg{criterion = 'MSE', real_label = 0.001, gen_label = 0, pixel_label = true,
                         pixel_div = opt.fineSize * opt.fineSize, lr = opt.lr, beta1 = opt.beta1}

local function rgb2gray(im)                                                   -- train-gray.lua:54-76
   local g = torch.FloatTensor(im:size(1), 1, im:size(3), im:size(4))
   g[{{}, 1}]:copy(im[{{}, 1}]):mul(0.21):add(0.72, im[{{}, 2}]):add(0.07, im[{{}, 3}])
   return g
end

local tm = torch.Timer()
for epoch = 1, opt.niter do
   for i = 1, math.min(data:size(), opt.ntrain), opt.batchSize do
      tm:reset()
      local real = rgb2gray(data:getBatch())                                  -- fDx's data:getBatch() (:213)
      local errD_real, errD_fake, errG = dsr.train_step(ctx, netG, netD, step, real)   -- optim.adam(fDx) + optim.adam(fGx) (:280-283)
      print(('Epoch: [%d][%8d / %8d]\t Time: %.3f  Err_G: %.4f  Err_D: %.4f'):format(
         epoch, (i - 1) / opt.batchSize, math.floor(math.min(data:size(), opt.ntrain) / opt.batchSize), tm:time().real,
         errG, errD_real + errD_fake))
   end
end
