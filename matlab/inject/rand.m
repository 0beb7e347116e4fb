function r = rand(varargin)
% RAND  Tape-replaying shadow of the built-in (see README.md in this directory).
global EMB_TAPE EMB_TAPE_POS
if isempty(EMB_TAPE), r = builtin('rand', varargin{:}); return; end
if nargin == 0, sz = [1 1]; elseif nargin == 1 && numel(varargin{1}) > 1, sz = varargin{1};
elseif nargin == 1, sz = [varargin{1} varargin{1}]; else, sz = [varargin{:}]; end
n = prod(sz);
r = reshape(EMB_TAPE(EMB_TAPE_POS + (1:n)), sz);
EMB_TAPE_POS = EMB_TAPE_POS + n;
end
