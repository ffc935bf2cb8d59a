package sw ;

import org.apache.spark.api.java.function.PairFunction ;

import scala.Tuple2 ;
import scala.Tuple3 ;

import java.util.ArrayList ;
import java.util.Collections ;
import java.util.Comparator ;
import java.util.List ;

/**
 * The two bodies of the reference's sw.Distribution that change when the hot path moves to libswb200
 * (everything else of Distribution.java -- DistributeReference's I/O and reduction, NoDistribution,
 * CombineReadsToRef, the comparators -- stays as it is).  Java 1.8; NOT COMPILED in the build environment.
 *
 * <ul>
 * <li>{@link MapRef#call}: replaces Distribution.java:403-436.  One native call for the element's reference
 *     against ALL reads (1 x R pairs) instead of R calls of OptAlignments; works unchanged inside a Spark
 *     executor of any deploy mode.</li>
 * <li>{@link #mapRefsBatched}: replaces the per-file `sc.parallelize(list).mapToPair(new MapRef())` of
 *     Distribution.java:337-338 when driver and GPU share a JVM (local[N]): ONE native call per reference
 *     file (all references x all reads), then the same per-reference (key, value) tuples MapRef returns;
 *     hand them to `sc.parallelizePairs(...)` so that sortByKey / first / lookup (:341-352) run unchanged.
 *     With -Dswb.devices=0,1,...  the file is sharded over several GPUs inside the native call
 *     (swb_multi_*: the reference's partition point moved below the JNI boundary).</li>
 * </ul>
 */
public class Distribution
{
	/** total score wraps like the reference's `int totalScore +=` (Distribution.java:424) */
	static class MapRef implements PairFunction< Tuple3<String[],ArrayList<String>,Tuple2<int[],char[]>> , Integer , Tuple2<String[],ArrayList<Tuple2<Integer,String[]>>> >
	{
		@Override
		public Tuple2<Integer,Tuple2<String[],ArrayList<Tuple2<Integer,String[]>>>> call( Tuple3<String[],ArrayList<String>,Tuple2<int[],char[]>> tuple )
		{
			String[] ref = tuple._1() ;
			ArrayList<String> reads = tuple._2() ;
			int[] alignScores = tuple._3()._1() ;
			List<Tuple2<Integer,ArrayList<Tuple2<Integer,String[]>>>> row =
					SmithWaterman.alignAll( Collections.singletonList( ref[1] ) , reads , alignScores , 0 ).get(0) ;
			return reduceRow( ref , row ) ;
		}
	}

	/** reads in file order, cells in list order, then the reference's stable sort by beginning (Distribution.java:419-428) */
	static Tuple2<Integer,Tuple2<String[],ArrayList<Tuple2<Integer,String[]>>>> reduceRow( String[] ref , List<Tuple2<Integer,ArrayList<Tuple2<Integer,String[]>>>> row )
	{
		int totalScore = 0 ;
		ArrayList<Tuple2<Integer,String[]>> matchSites = new ArrayList<Tuple2<Integer,String[]>>() ;
		for( Tuple2<Integer,ArrayList<Tuple2<Integer,String[]>>> result : row )
		{
			totalScore += result._1().intValue() ;
			matchSites.addAll( result._2() ) ;
		}
		Collections.sort( matchSites , new Comparator<Tuple2<Integer,String[]>>()       // == MatchSiteComp (Distribution.java:691-694)
		{
			public int compare( Tuple2<Integer,String[]> a , Tuple2<Integer,String[]> b ) { return a._1().intValue() - b._1().intValue() ; }
		} ) ;
		Tuple2<String[],ArrayList<Tuple2<Integer,String[]>>> value = new Tuple2<String[],ArrayList<Tuple2<Integer,String[]>>>( ref , matchSites ) ;
		return new Tuple2<Integer,Tuple2<String[],ArrayList<Tuple2<Integer,String[]>>>>( Integer.valueOf( totalScore ) , value ) ;
	}

	/** one native call for a whole reference file: element r == MapRef.call( (refSeqs.get(r), reads, algoArgs) ) */
	static ArrayList<Tuple2<Integer,Tuple2<String[],ArrayList<Tuple2<Integer,String[]>>>>> mapRefsBatched( ArrayList<String[]> refSeqs , ArrayList<String> reads , Tuple2<int[],char[]> algoArgs )
	{
		ArrayList<String> seqs = new ArrayList<String>( refSeqs.size() ) ;
		for( String[] r : refSeqs ) seqs.add( r[1] ) ;
		List<List<Tuple2<Integer,ArrayList<Tuple2<Integer,String[]>>>>> all = SmithWaterman.alignAll( seqs , reads , algoArgs._1() , 0 ) ;
		ArrayList<Tuple2<Integer,Tuple2<String[],ArrayList<Tuple2<Integer,String[]>>>>> out = new ArrayList<Tuple2<Integer,Tuple2<String[],ArrayList<Tuple2<Integer,String[]>>>>>( refSeqs.size() ) ;
		for( int r = 0 ; r < refSeqs.size() ; r++ ) out.add( reduceRow( refSeqs.get(r) , all.get(r) ) ) ;
		return out ;
	}
}
