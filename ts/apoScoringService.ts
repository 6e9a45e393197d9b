/*---------------------------------------------------------------------------------------------
 *  NEW FILE for src/vs/workbench/contrib/senweaver/common/apoScoringService.ts
 *
 *  Common-layer contract + renderer-side proxy of the B200 scoring engine.  The `common`
 *  layer may not load native code (eslint.config.js:94-111), so this file only declares the
 *  service and builds a ProxyChannel client; the implementation lives in the main process
 *  (apoScoringMainService.ts), exactly like the metrics service of the reference
 *  (common/metricsService.ts:25-50 / electron-main/metricsMainService.ts:35).
 *
 *  This file cannot be compiled in the engine's build image (no tsc / node); it is the
 *  reference-side binding INTEGRATION.md describes.
 *--------------------------------------------------------------------------------------------*/
import { VSBuffer } from '../../../../base/common/buffer.js';
import { ProxyChannel } from '../../../../base/parts/ipc/common/ipc.js';
import { createDecorator } from '../../../../platform/instantiation/common/instantiation.js';
import { registerSingleton, InstantiationType } from '../../../../platform/instantiation/common/extensions.js';
import { IMainProcessService } from '../../../../platform/ipc/common/mainProcessService.js';

export const APO_SCORING_CHANNEL = 'senweaver-channel-apoScoring';

/** Raw result blocks; layouts are those of include/apo_b200.h (see traceRecordCodec.ts). */
export interface ApoScoreBlocks {
	scores: VSBuffer;   // float64[C]  (-Infinity: candidate without a non-null evaluation)
	counts: VSBuffer;   // uint64[C]
	topk: VSBuffer;     // int32[K]    score desc, ties -> lower index
	report: VSBuffer;   // struct apo_corpus_report (784 bytes)
}

export interface IApoScoringService {
	readonly _serviceBrand: undefined;
	/** false when the addon or a B200 is unavailable: callers keep their state untouched and warn. */
	isAvailable(): Promise<boolean>;
	/** TraceCollectorService._computeRewardSignals for n packed records (apo_reward_batch). */
	rewardBatch(records: VSBuffer): Promise<{ dims: VSBuffer; masks: VSBuffer; finals: VSBuffer }>;
	/** The string stored under 'senweaver.traceCollector.data' -> packed Form R records (apo_records_from_json):
	 *  lets a stored corpus reach the engine without materialising ConversationTrace objects. */
	recordsFromJson(persisted: string): Promise<VSBuffer>;
	/** Corpus report (APOService._buildReport numeric content) + candidate scores / top-K.
	 *  dims: float32[C][T][9] with NaN = dimension absent (pass C = 1, T = 4 of NaN for report-only calls). */
	score(dims: VSBuffer, C: number, T: number, corpus: VSBuffer | undefined, K: number): Promise<ApoScoreBlocks>;
}

export const IApoScoringService = createDecorator<IApoScoringService>('senweaverApoScoringService');

/** Renderer-side client: everything that crosses the channel is async (metricsService.ts:47). */
export class ApoScoringService implements IApoScoringService {
	readonly _serviceBrand: undefined;
	private readonly _proxy: IApoScoringService;

	constructor(@IMainProcessService mainProcessService: IMainProcessService) {
		this._proxy = ProxyChannel.toService<IApoScoringService>(mainProcessService.getChannel(APO_SCORING_CHANNEL));
	}

	isAvailable(): Promise<boolean> { return this._proxy.isAvailable(); }
	rewardBatch(records: VSBuffer) { return this._proxy.rewardBatch(records); }
	recordsFromJson(persisted: string) { return this._proxy.recordsFromJson(persisted); }
	score(dims: VSBuffer, C: number, T: number, corpus: VSBuffer | undefined, K: number) { return this._proxy.score(dims, C, T, corpus, K); }
}

registerSingleton(IApoScoringService, ApoScoringService, InstantiationType.Delayed);
